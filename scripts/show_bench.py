"""print the headline numbers of bench JSON lines (development aid)"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    r, g = d.get("roofline") or {}, d.get("gcfm") or {}
    print(f"{f}: N={d.get('n_gpus')} value {d.get('value'):.1f} {d.get('unit')} e2e {(d.get('e2e') or {}).get('value', 0):.1f} "
          f"ms/step {d.get('ms_per_step'):.2f} roof {r.get('frac', 0):.3f} (launch {r.get('avg_launch_ms', 0):.4f} ms, share {r.get('share_of_step') or 0:.3f}) "
          f"gcfm {g.get('value', 0)/1e6:.2f}M dev / {g.get('e2e_value', 0)/1e6:.2f}M e2e  parity {d.get('parity')}")
