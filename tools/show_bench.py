#!/usr/bin/env python3
"""Condensed view of a bench.py JSON line (file argument)."""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f"N={d['n_gpus']} value {d['value']:.1f} {d['unit']}  {d['ms_per_step']:.1f} ms/step  launches {d['gpu_launches']}")
print("  phases", {k: round(v, 2) for k, v in d["phases_ms"].items()})
e = d.get("e2e")
if e:
    print(f"  e2e {e['value']:.1f}  {e['ms_per_step']:.1f} ms  h2d {e['h2d_bytes_per_step']} d2h {e['d2h_bytes_per_step']}", {k: round(v, 2) for k, v in e["phases_ms"].items()})
if d.get("parity"):
    print("  parity", {k: v for k, v in d["parity"].items() if k != "how"})
r = d["roofline"]
print(f"  roofline frac {r['frac']:.3f} (mix {r['frac_of_mix_probe']:.3f}) kernel {r['kernel_ms']:.1f} ms  {r['gpairs_evaluated_per_s_kernel']:.1f} Gpairs/s evaluated  traffic {r['traffic']}")
for c in (d.get("eps") or {}).get("cases", []):
    rs, rc = c.get("roofline_sweep") or {}, c.get("roofline_csr") or {}
    print(f"  eps {c['name']}: {c['ms_per_build']:.2f} ms {c['gpairs_per_s']:.1f} Gp/s nnz {c['nnz']} [{c['path']}] sweep frac {rs.get('frac', 0):.3f} csr frac {rc.get('frac', 0):.3f}"
          f" parity {c.get('parity', {}).get('ok')} e2e {(c.get('e2e') or {}).get('ms_per_build')}", {k: round(v, 2) for k, v in c["phases_ms"].items()})
for c in (d.get("queries") or {}).get("cases", []):
    print(f"  query {c['name']}: {c['ms']:.2f} ms {c['gpairs_per_s']:.1f} Gp/s", [round(v, 1) for v in c.get("ms_each", [])])
print("  clocks", d.get("clocks"))
if d.get("cpu_baseline"):
    print("  cpu", d["cpu_baseline"]["value"], (d["cpu_baseline"].get("best_effort_c") or {}).get("value"))
