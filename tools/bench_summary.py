#!/usr/bin/env python
"""One-screen summary of bench.py JSON lines: python tools/bench_summary.py gpurun_out/x.json ..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as exc:  # noqa: BLE001
        print(f, "unreadable:", exc)
        continue
    us = lambda x: x["ms_per_step"] * 1e3  # noqa: E731
    if d.get("impl") == "reference":
        print(f"{f}: reference arm {d['value'] / 1e9:.3f} G on {d['cpu_baseline']['cores']} cores  ({d['cpu_baseline']['sample']})")
        continue
    print(f"{f}: n_gpus={d['n_gpus']} value {d['value'] / 1e9:.1f} G  {us(d):.3f} us/step  frac {d['roofline']['frac']:.3f}  "
          f"clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
    ssd = d.get("single_stream_default")
    if ssd:
        print(f"   one stream pdl=1: cold {us(ssd['cold_ring']):.3f} us ({ssd['cold_ring']['frac']:.3f})  resident {us(ssd['l2_resident']):.3f} us")
    print(f"   one stream pdl=2: cold {us(d['single_stream_chained']):.3f} us ({d['single_stream_chained']['frac']:.3f})  "
          f"resident {us(d['l2_resident']):.3f} us")
    w = d.get("single_stream_wide")
    if w:
        print(f"   one stream, wide build: pdl=1 cold {us(w['default_cold_ring']):.3f} us ({w['default_cold_ring']['frac']:.3f})  "
              f"pdl=2 cold {us(w['chained_cold_ring']):.3f} us ({w['chained_cold_ring']['frac']:.3f})")
    r = d.get("rollout")
    if r:
        print(f"   rollout {r['value'] / 1e9:.1f} G  write {r['write_gbs']:.0f} GB/s = {r.get('frac_of_write_peak', 0):.3f} of fill "
              f"{r['hbm_write_only_gbs_measured']:.0f}; {r.get('frac_of_copy_peak', r.get('frac', 0)):.3f} of copy peak")
    e = d.get("e2e")
    if e:
        print(f"   e2e {e['value'] / 1e9:.2f} G  {e['pcie_gbs']:.1f} GB/s of ceiling {e.get('pcie_ceiling_gbs', 0):.1f} "
              f"({e.get('frac_of_pcie', 0):.3f})  sync {e['synchronous_value'] / 1e9:.2f} G  compact "
              f"{e.get('compact', {}).get('value', 0) / 1e9:.2f} G")
    if d.get("cpu_baseline"):
        print(f"   cpu {d['cpu_baseline']['value'] / 1e9:.3f} G on {d['cpu_baseline']['cores']} cores")
