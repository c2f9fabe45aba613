"""Print the parity / timing summary of a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
def show(name, p):
    print(name, "ok", p["ok"], "lat", p.get("latents_checked"))
    if "per_tensor_rank0" not in p:
        print("   gross (max over ranks)", {k: "%.1e" % v for k, v in p["all_latents_vs_torch_cuda_oracle"]["per_tensor_max_over_ranks"].items()})
    for k in p.get("per_tensor_rank0", {}):
        ve = p["vs_exact_rank0"]
        vd = ve.get("reference_torch_cuda", {})
        print("   %-8s err %.2e | exact: ours %s host %s cuda %s | floor %.2e allow %.2e | gross %.2e | %s" % (
            k, p["per_tensor_rank0"][k], "%.2e" % ve["ours"][k] if k in ve["ours"] else "   -    ",
            "%.2e" % ve["reference"][k] if k in ve["reference"] else "   -    ", "%.2e" % vd[k] if k in vd else "   -    ", p["input_rounding_floor_rank0"][k],
            p["allowance_per_tensor_rank0"][k], p["all_latents_vs_torch_cuda_oracle"]["per_tensor_rank0"][k], p["verdict_rank0"][k]))
    if "cross_rank" in p:
        cr = p["cross_rank"]
        print("   cross_rank", cr["ok"], "bit-identical ranks:", cr["ranks_bit_identical"], {k: "%.1e/%.1e" % (v, cr.get("tol_per_tensor_rank0", {}).get(k, 0)) for k, v in cr["per_tensor_rank0"].items()})
if d.get("parity"):
    show(d["config"]["workload"][:4], d["parity"])
for n, v in d.get("other_configs", {}).items():
    if v.get("parity"): show(n, v["parity"])
    elif "failed" in v: print(n, v)
print("headline %.0f subj/s  %.3f ms/step  e2e %.0f  plain %.0f  frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["plain_loop_value"], d["roofline"]["frac"]), d["roofline"]["phase_ms"])
for n, v in d.get("other_configs", {}).items():
    if "value" in v:
        print("%s %.0f subj/s %.3f ms  e2e %.0f  frac %.3f" % (n, v["value"], v["ms_per_step"], v["e2e"]["value"], v["roofline_frac"] or 0), {k: round(x, 3) for k, x in v["phase_ms"].items()}, v["kernel_path"][:40])
print("cpu_baseline", d.get("cpu_baseline"), "\ngpu_reference", d.get("gpu_reference"))
