"""Summarise the source page of an ncu report: stall reasons per SASS opcode and the instructions with most samples.
usage: python scripts/ncu_source.py rep.ncu-rep [kernel-index] [--row N]   (development aid)"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = None, []
for r in rows:
    if r and r[0] == "Address":
        if hdr is None:
            hdr = r
        else:
            break
    elif hdr and len(r) == len(hdr):
        data.append(r)
idx = {h: i for i, h in enumerate(hdr)}
S = lambda r, c: int(r[idx[c]] or 0)
tot = sum(S(r, "# Samples") for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = collections.Counter()
byop, cnt = collections.Counter(), collections.Counter()
for r in data:
    src = r[idx["Source"]].split()
    op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
    byop[op] += S(r, "# Samples"); cnt[op] += 1
    for c in stall_cols:
        st[c] += S(r, c)
print("total samples", tot, "instructions", len(data))
for op, s in byop.most_common(12):
    print(f"{op:10s} {s:7d} {100*s/tot:5.1f}%  n={cnt[op]:4d}  per-inst {s/cnt[op]:.1f}")
print({k: round(100 * v / tot, 1) for k, v in st.most_common(12)})
if "--row" in sys.argv:
    n = int(sys.argv[sys.argv.index("--row") + 1])
    bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[idx["Source"]]]
    a, b = bars[n], bars[n + 1]
    print(f"row between barriers {n} and {n+1}: {b-a} instructions, {sum(S(r,'# Samples') for r in data[a:b])} samples")
    for r in data[a:b + 1]:
        top = sorted(((S(r, c), c[6:]) for c in stall_cols if c != "stall_selected"), reverse=True)[:2]
        print(f"{S(r,'# Samples'):4d} sel{S(r,'stall_selected'):3d} " + " ".join(f"{c}={v}" for v, c in top if v) .ljust(34) + " | " + r[idx["Source"]].strip()[:80])
