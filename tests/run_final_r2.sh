ncu --clock-control none --set full --import-source on -k regex:qsgd_encode_chunks -s 1 -c 1 -f -o gpurun_out/prof_qsgd_enc_r2 python bench.py --codec qsgd --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_qsgd_enc_r2.log 2>&1
ncu -i gpurun_out/prof_qsgd_enc_r2.ncu-rep --page raw --csv > gpurun_out/prof_qsgd_enc_r2_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_qsgd_enc_r2.ncu-rep --page source --csv --print-source sass > /tmp/src_q.csv 2>/dev/null
python profiles/sass_hot.py /tmp/src_q.csv > gpurun_out/prof_qsgd_enc_r2_hot.txt 2>&1
rm -f gpurun_out/prof_qsgd_enc_r2.ncu-rep
ncu --clock-control none --set full --import-source on -k regex:sign_decode_reduce -s 3 -c 1 -f -o gpurun_out/prof_sign_dec2_r2 python bench.py --codec sign --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu -i gpurun_out/prof_sign_dec2_r2.ncu-rep --page raw --csv > gpurun_out/prof_sign_dec2_r2_raw.csv 2>/dev/null
rm -f gpurun_out/prof_sign_dec2_r2.ncu-rep
bash tests/run_profile_r2.sh r2h quick > /dev/null 2>&1
for c in qsgd terngrad sign topk; do timeout 600 python bench.py --codec $c --steps 50 --warmup 5 > gpurun_out/bench_${c}_r2h.json 2> gpurun_out/bench_${c}_r2h.err; done
timeout 600 python bench.py --codec hsq --k-bit 12 --workload flat --steps 10 --warmup 3 > gpurun_out/bench_hsq_k12_r2h.json 2>/dev/null
timeout 600 python bench.py --codec hsq --c-dim 8 --workload flat --steps 20 --warmup 3 > gpurun_out/bench_hsq_d8_r2h.json 2>/dev/null
timeout 600 python bench.py --codec hsq --c-dim 32 --workload flat --steps 20 --warmup 3 > gpurun_out/bench_hsq_d32_r2h.json 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*_r2h.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d["roofline"]; c = d.get("cpu_baseline") or {}
        print("%-34s step %.1f us %.1f Gelem/s enc %.1f (frac %.3f) dec %.1f (frac %.3f) e2e %.2f ms cpu port %.3f ref %.4f Gelem/s" % (
            f.split("/")[-1], d["ms_per_step"]*1e3, d["value"]/1e9, r["encode_ms"]*1e3, r["frac"], r["decode_ms"]*1e3, r["decode_frac"],
            d["e2e"]["ms_per_step"], c.get("value", 0)/1e9, c.get("reference_py_value", 0)/1e9))
    except Exception as e:
        print(f, "parse failed", e)
PY
