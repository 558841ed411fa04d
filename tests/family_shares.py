"""Developer tool: kernel-family shares of the step from an ncu launch list (gpu__time_duration.sum csv of tests/profile_step.py)
-> profiles/family_shares.json (read by bench.py).  python tests/family_shares.py profiles/r02_launches_650.csv"""
import collections
import csv
import json
import re
import sys

FAMILIES = [
    ("gemm_stream_kernel (streaming tcgen05 GEMM: HBM-bound 1x1 convs of layer1-3, stem, FFN1)", r"gemm_stream_kernel"),
    ("conv3x3_halo_kernel (halo-reusing 3x3 / 64-channel conv, forward + data gradient)", r"conv3x3_halo"),
    ("gemm_tcp_kernel (persistent tcgen05 GEMM / conv: tensor-bound shapes)", r"gemm_tcp_kernel|gemm_pair_kernel"),
    ("gemm_tc_kernel (one-tile tcgen05 GEMM / conv: small transformer linears, strided convs and data gradients)", r"gemm_tc_kernel"),
    ("wgrad_tc_kernel (tcgen05 weight gradients, side stream)", r"wgrad_tc_kernel|wgrad_narrow_kernel"),
    ("attention forward / backward (tcgen05)", r"attn_"),
    ("LayerNorm forward / backward", r"ln_(fwd|bwd)_kernel"),
    ("Adam + clipnorm + weight refresh", r"chunk_|adam_|prep_weights"),
    ("pooling / input layout / stride-2 scatter / fills", r"maxpool|image_|s2d|add_rowbcast|scatter_s2|FillFunctor|elementwise"),
    ("matcher + set loss", r"matcher_kernel|set_loss"),
    ("mma.sync GEMM / wgrad (unaligned heads)", r"igemm_kernel|wgrad_kernel"),
]
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
acc = collections.OrderedDict((name, [0, 0.0]) for name, _ in FAMILIES)
other, total = [0, 0.0], 0.0
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    try:
        us = float(r[vi].replace(",", "")) / 1e3
    except ValueError:
        continue
    total += us
    for name, pat in FAMILIES:
        if re.search(pat, r[ki]):
            acc[name][0] += 1
            acc[name][1] += us
            break
    else:
        other[0] += 1
        other[1] += us
out = {"source": f"{path} (ncu --metrics gpu__time_duration.sum --clock-control none, one eager step B=8 800x1333; serialised, cold cache: "
                 "shares, not absolutes)", "total_us": round(total, 1),
       "families": {k: {"launches": v[0], "us": round(v[1], 1), "share": round(v[1] / total, 4)} for k, v in acc.items() if v[0]}}
if other[0]:
    out["families"]["other"] = {"launches": other[0], "us": round(other[1], 1), "share": round(other[1] / total, 4)}
json.dump(out, open("profiles/family_shares.json", "w"), indent=1)
for k, v in sorted(out["families"].items(), key=lambda kv: -kv[1]["us"]):
    print(f"{v['us']:9.1f} us {100 * v['share']:5.1f}% n={v['launches']:4d}  {k}")
