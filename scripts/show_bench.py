"""Pretty-prints a bench.py JSON line: headline, e2e variants, roofline and the per-config table.  usage: python scripts/show_bench.py gpurun_out/bench_rXX.json"""
import json, sys
l = json.loads(open(sys.argv[1]).read())
print(f"value {l['value']:.0f} {l['unit']} ms/step {l['ms_per_step']:.3f} launches {l['gpu_launches']} clocks {l['clocks']}")
e = l["e2e"]
print("e2e", {k: (round(v) if isinstance(v, float) else v) for k, v in e.items() if k in ("value", "pageable", "pageable_registered", "pcie_ceiling_fps", "in_flight", "pageable_in_flight")})
r = l["roofline"]; print("roofline", r["kernel"], f"{r['achieved']:.0f}/{r['peak']:.0f} = {r['frac']:.3f} traffic {r['traffic']}")
print("path", {k: round(v, 3) for k, v in l["path"].items()})
print("cpu", l["cpu_baseline"] and (round(l["cpu_baseline"]["value"]), l["cpu_baseline"]["cores"]))
def show(d, ind=0):
    for k, v in d.items():
        if isinstance(v, dict):
            if "fps" in v:
                print(" " * ind + f"{k}: fps={v['fps']:.0f} us/frame={v.get('us_per_frame_per_gpu', 0):.2f} frac={v.get('frac_of_hbm_peak', 0):.3f} {v.get('one_read', '')}")
            else:
                print(" " * ind + k + ":")
            show({kk: vv for kk, vv in v.items() if isinstance(vv, dict)}, ind + 2)
if l.get("configs"):
    show(l["configs"])
