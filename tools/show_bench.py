import json,sys
d=json.load(open(sys.argv[1]))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["ms_each_step"], d["infer"]["value"] if d.get("infer") else None)
for k,v in d["kernels"].items(): print("%-24s %4d %7.3f %s"%(k,v["launches"]//d["steps"],v["ms_per_step"],v["tflops"]))
