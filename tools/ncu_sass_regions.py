"""Per-SASS-instruction counters of one kernel in an .ncu-rep, summed over address ranges.
usage: python tools/ncu_sass_regions.py x.ncu-rep <kernel-substring> [--dump] [--regions name:lo-hi,...]   (hex offsets from kernel start)"""
import csv, io, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kern, hdr, data = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name": kern = r[1]; hdr = None; continue
    if r[0] == "Address": hdr = r; continue
    if hdr and kern and want in kern and len(r) == len(hdr): data.append(r)
iE, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = int(data[0][0], 16)
ins = [(int(r[0], 16) - base, r[1].strip(), int(r[iE]), int(r[iT]), int(r[iS])) for r in data]
totE = sum(i[2] for i in ins); totS = sum(i[4] for i in ins)
print("kernel %s: %d SASS, %d warp-inst, %d stall samples" % (want, len(ins), totE, totS))
if "--dump" in sys.argv:
    for off, s, e, t, smp in ins:
        print("%05x %10d %5.1f %6d  %s" % (off, e, t / max(e, 1), smp, s))
if "--regions" in sys.argv:
    for g in sys.argv[sys.argv.index("--regions") + 1].split(","):
        nm, rng = g.split(":"); lo, hi = [int(x, 16) for x in rng.split("-")]
        sel = [i for i in ins if lo <= i[0] <= hi]
        e = sum(i[2] for i in sel); t = sum(i[3] for i in sel); s = sum(i[4] for i in sel)
        print("  %-24s %4d SASS %6.2f%% exec %6.2f%% stall-samples %5.1f lanes" % (nm, len(sel), 100 * e / totE, 100 * s / totS, t / max(e, 1)))
