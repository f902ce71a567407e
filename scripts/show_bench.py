import json, sys
d = json.load(open(sys.argv[1]))
for k in ("value", "ms_per_step", "e2e", "gpu_launches", "model_flops_utilisation", "profiled_kernel_ms_per_step", "peak_mem_gb", "clocks", "cpu_baseline", "loss"):
    if k in d:
        print(k, ":", d[k])
r = d.get("roofline") or {}
print("roofline:", {k: r.get(k) for k in ("kernel", "achieved", "frac", "share_of_step")})
for k, v in sorted((d.get("kernels") or {}).items(), key=lambda kv: -kv[1]["ms"]):
    print(f"   {k:20s} n={v['launches']:4d} ms={v['ms']:8.3f}  " + (f"tflops={v['tflops']:.0f} ({v['frac_tensor_peak']:.2f})" if 'tflops' in v else f"GB/s={v.get('gbs', 0):.0f} ({v.get('frac_hbm_peak', 0):.2f})"))
s = d.get("sample")
if s:
    print("sample:", {k: s.get(k) for k in ("value", "ms_per_step", "steps", "finite")}, "e2e", s.get("e2e", {}).get("value"),
          "frac", (s.get("roofline") or {}).get("frac"))
