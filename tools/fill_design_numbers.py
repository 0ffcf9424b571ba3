"""Here (no GPU): fill the R2_* placeholders of DESIGN.md from the bench lines under profiles/.
Usage: python tools/fill_design_numbers.py   (idempotent once the placeholders are gone)"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def line(name):
    with open(os.path.join(P, name)) as f:
        rows = [l for l in f if l.startswith("{")]
    return json.loads(rows[-1])


n1, n2, n8, ref = line("r2_bench_n1.json"), line("r2_bench_n2.json"), line("r2_bench_n8.json"), line("r2_bench_reference_arm.json")
sweep = {b["batch_per_gpu"]: b for b in n1["batch_sweep"]}
fmt = lambda v: f"{v:,.0f}".replace(",", " ")
vals = {
    "R2_MS": f"{n1['ms_per_step']:.3f}",
    "R2_VALUE": fmt(n1["value"]),
    "R2_E2E": fmt(n1["e2e"]["value"]),
    "R2_REF": f"{ref['value']:.1f}",
    "R2_N2_EFF": f"{n2['value'] / (2 * n2['single_gpu_same_run']['value']):.3f}",
    "R2_N2_COS": f"{n2['dp_parity']['grad_cos']:.5f}",
    "R2_N2": fmt(n2["value"]),
    "R2_N8_EFF": f"{n8['value'] / (8 * n8['single_gpu_same_run']['value']):.3f}",
    "R2_N8_X": f"{n8['value'] / n8['single_gpu_same_run']['value']:.2f}",
    "R2_N8_E2E": fmt(n8["e2e"]["value"]),
    "R2_N8": fmt(n8["value"]),
    "R2_CV2": fmt(n2["cv1"]["value"]),
    "R2_CV8": fmt(n8["cv1"]["value"]),
    "R2_CV_COS": f"{n8['cv1']['dp_parity']['grad_cos']:.5f}",
    "R2_B8": fmt(sweep[8]["patches_per_sec"]),
    "R2_B32": fmt(sweep[32]["patches_per_sec"]),
    "R2_B128": fmt(sweep[128]["patches_per_sec"]),
    "R2_INF_E2E": f"{n1['inference']['e2e_value'] / 1e3:.1f}",
    "R2_INF": f"{n1['inference']['value'] / 1e3:.1f}",
}
path = os.path.join(ROOT, "DESIGN.md")
s = open(path).read()
for k in sorted(vals, key=len, reverse=True):          # longest first: R2_N2_EFF before R2_N2
    s = re.sub(r"\b" + k + r"\b", vals[k], s)
open(path, "w").write(s)
left = sorted(set(re.findall(r"\bR2_[A-Z0-9_]+\b", s)))
print("filled", len(vals), "placeholders; left:", left)
