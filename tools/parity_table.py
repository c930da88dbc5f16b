"""Render profiles/parity_r02.txt from gpurun_out/parity_r02.jsonl (written by tests/test_gpu_fullsize_golden.py on the
B200 box) plus the driver-visible bench line's parity_rel_err."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out", "parity_r02.jsonl")
rows = {}
for line in open(src):
    r = json.loads(line)
    rows[r["fixture"]] = r          # last run of a fixture wins
out = ["Parity of the decode kernel at BASELINE geometry, measured on the B200 box (tests/test_gpu_fullsize_golden.py).",
       "Reference outputs: the UNMODIFIED reference run on CPU (oracle/make_golden_fullsize.py).  rel = max|a-b| / max|b|.",
       "ref noise = the reference's own torch operators replayed on the B200 in fp16 (oracle/transformer_torch.py) against the",
       "            reference's CPU fp16 output: the distance between two legitimate fp16 executions of the same path.",
       "",
       f"{'fixture':12s} {'ours vs ref fp16':>17s} {'ref fp16 noise':>15s} {'ours vs ref fp32':>17s} {'ref fp16 vs fp32':>17s} {'worst probe':>12s}"]
for k in sorted(rows):
    r = rows[k]
    out.append(f"{k:12s} {r['ours_vs_ref_fp16']:17.2e} {r['ref_fp16_order_noise']:15.2e} {r['ours_vs_ref_fp32']:17.2e} "
               f"{r['ref_fp16_vs_ref_fp32']:17.2e} {r['worst_probe']:12d}")
out.append("")
for k in sorted(rows):
    r = rows[k]
    out.append(f"{k}: per probe position vs ref fp16: " +
               ", ".join(f"{p}:{e:.1e}" for p, e in zip(r["probes"], r["per_probe_vs_ref_fp16"])))
text = "\n".join(out) + "\n"
dst = os.path.join(ROOT, "profiles", "parity_r02.txt")
open(dst, "w").write(text)
sys.stdout.write(text)
