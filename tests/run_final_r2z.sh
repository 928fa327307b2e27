#!/bin/bash
# Final round-2 session on one B200: test-suite, one bench line per codec / config, launch list, ncu captures.
TAG=r2z
bash tests/run_profile_r2.sh $TAG
rm -f gpurun_out/prof_enc_$TAG.ncu-rep gpurun_out/prof_dec_$TAG.ncu-rep
bash tests/run_ncu_r2.sh $TAG
for w in t5; do timeout 600 python bench.py --codec sign --sign-wire $w --steps 50 --warmup 5 > gpurun_out/bench_sign_${w}_$TAG.json 2>/dev/null; done
rm -f gpurun_out/prof_enc_$TAG.ncu-rep
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*_r2z.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d["roofline"]; c = d.get("cpu_baseline") or {}
        print("%-34s step %.1f us %.1f Gelem/s enc %.1f (frac %.3f) dec %.1f (frac %.3f) e2e %.2f ms cpu port %.3f ref %.4f Gelem/s" % (
            f.split("/")[-1], d["ms_per_step"]*1e3, d["value"]/1e9, r["encode_ms"]*1e3, r["frac"], r["decode_ms"]*1e3, r["decode_frac"],
            d["e2e"]["ms_per_step"], c.get("value", 0)/1e9, c.get("reference_py_value", 0)/1e9))
    except Exception as e:
        print(f, "parse failed", e)
PY
