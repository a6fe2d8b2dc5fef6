"""Print the headline fields of a bench.py JSON line."""
import json
import sys

b = json.loads([l for l in open(sys.argv[1]) if l.lstrip().startswith("{")][-1])   # the JSON line (torchrun may print banners)
print("value %.3f %s  ms/step %.2f  e2e %.3f" % (b["value"], b["unit"], b["ms_per_step"], b["e2e"]["value"]))
print("kernel ms/step", {k: round(v, 2) for k, v in b["kernel_ms_per_step"].items()})
print("objective %.10f grad_norm %.10f launches %s" % (b["objective"], b["grad_norm"], b.get("gpu_launches")))
