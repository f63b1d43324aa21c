import json,sys
d=json.load(open(sys.argv[1]))
for k,v in d.items():
    if isinstance(v,dict):
        print(k, round(v["ms_per_step"],4), "ms", list(v["layer1_phase_ns"]["detail_rel_to_P1_start"].values()))
        pr=v.get("prof_cycles_mbar_cnt_total")
        if pr:
            import statistics as st
            for j,n in enumerate(("mbar","cnt","total")):
                xs=[p[j] for p in pr]
                print("   ", n, "min", min(xs), "med", int(st.median(xs)), "max", max(xs))
            print("    per-cta mbar:", [p[0]//1000 for p in pr][:148:4])
