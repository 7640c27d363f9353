"""Per-region breakdown of one kernel of an .ncu-rep: runs of SASS instructions with a similar execution count, with their share of
the executed instructions (per `--unit` warp-level work items), stall samples and shared-memory wavefronts.
    python tools/ncu_regions.py rep.ncu-rep k_ring_partition 31.25e6"""
import csv, math, subprocess, sys
rep, kern, unit = sys.argv[1], sys.argv[2], float(sys.argv[3])
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(r for r in rows if 'Instructions Executed' in r)
iex, isamp, iw, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared"), hdr.index("Source")
data, seen = [], set()
for r in rows:
    if len(r) == len(hdr) and r[iex].isdigit():
        if r[0] in seen:  # a second launch of the same kernel: keep the first only
            break
        seen.add(r[0])
        data.append(r)
out = [(int(r[iex]), int(r[isamp]), r[isrc].strip()[:70], int(r[iw] or 0)) for r in data]
tot, tsamp = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"# {kern}: {len(out)} SASS instructions, {tot} executed = {tot / unit:.2f} per unit, {tsamp} samples")
def flush(a, b):
    ex, sm, wf = sum(o[0] for o in out[a:b]), sum(o[1] for o in out[a:b]), sum(o[3] for o in out[a:b])
    if ex / unit > 0.05 or sm / max(tsamp, 1) > 0.003:
        print(f"{a:5d}-{b:5d} n={b - a:4d} exec={ex / unit:7.2f}/unit samples={100 * sm / tsamp:5.1f}% smem_wavefronts={wf / unit:5.2f}/unit  first: {out[a][2]}")
prev, seg = None, 0
for i, o in enumerate(out):
    lvl = -99 if o[0] == 0 else round(math.log(o[0], 1.5))
    if prev is None:
        prev = lvl
    if lvl != prev:
        flush(seg, i)
        seg, prev = i, lvl
flush(seg, len(out))
