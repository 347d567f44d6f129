import json, sys
for f in [x for x in sys.argv[1:] if x != "-v"]:
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line)
            print(f, "value %.1f e2e %.1f ms/step %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "roofline frac %.3f exec %.0f TF"%(d["roofline"]["frac"], d["roofline"]["executed_mma_tflops"]))
            print("  stages", {k[:28]:round(v,3) for k,v in d["stage_ms_per_step"].items()})
            if "layers" in d and "-v" in sys.argv:
                for l in d["layers"]:
                    print("    %-44s %8.3f ms  %7.1f alg TF" % (l["kernel"], l["ms"], l["alg_tflops"]))
