"""Hot source lines of one kernel in an ncu report (--import-source on): samples per CUDA source line with the top stall
reasons.  python scratch/ncu_source_hot.py <rep> <kernel regex> [top N]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fpath, hdr, agg, tot = None, None, collections.OrderedDict(), 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    if r[0] == "": continue                      # SASS line
    idx = {h: i for i, h in enumerate(hdr)}
    try: ns = int(r[idx["# Samples"]])
    except Exception: continue
    key = (fpath, int(r[0]), r[1].strip()[:110])
    st = {h[6:]: int(r[i] or 0) for h, i in idx.items() if h.startswith("stall_") and "(Not Issued)" not in h}
    inst = int(r[idx["Instructions Executed"]] or 0)
    a = agg.setdefault(key, [0, collections.Counter(), 0])
    a[0] += ns; a[1].update(st); a[2] += inst
    tot += ns
print(f"total samples {tot}")
for key, (ns, st, inst) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    reasons = ' '.join(f'{k}={v}' for k, v in st.most_common(3) if v)
    print(f"{ns:7d} {100*ns/max(tot,1):5.1f}%  inst={inst:9d}  {key[0]}:{key[1]:<4d} {key[2]}\n{'':16s}{reasons}")
