"""Per-CUDA-source-line cost of the kernels in an .ncu-rep (read here, no GPU).
usage: python tools/ncu_lines.py x.ncu-rep [kernel-substring] [top N]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fn, fpath, hdr = None, None, None
seen = set()
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue
    if want not in fn:
        continue
    key = (fn, fpath)
    iE, iS, iT = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    agg.setdefault(fn, []).append((fpath, int(r[0]), r[1].strip(), int(r[iE]), int(r[iS]), int(r[iT])))
done = set()
for fn, lines in agg.items():
    sig = (fn, sum(l[3] for l in lines))
    if sig in done:
        continue
    done.add(sig)
    totE = sum(l[3] for l in lines) or 1
    totS = sum(l[4] for l in lines) or 1
    print("== %s: %d warp-instructions, %d samples" % (fn[:100], totE, totS))
    for l in sorted(lines, key=lambda l: -l[3])[:top]:
        print("  %5.2f%% exec %5.2f%% smp %4.1f thr  %s:%d  %s" % (100 * l[3] / totE, 100 * l[4] / totS, l[5] / max(l[3], 1), l[0], l[1], l[2][:100]))

if "--groups" in sys.argv:
    # usage: ... --groups name:file:lo-hi,name:file:lo-hi ...   (sums per group, first matching kernel)
    spec = sys.argv[sys.argv.index("--groups") + 1]
    groups = []
    for g in spec.split(","):
        nm, f, rng = g.split(":")
        lo, hi = rng.split("-")
        groups.append((nm, f, int(lo), int(hi)))
    for fn, lines in agg.items():
        totE = sum(l[3] for l in lines) or 1
        acc = collections.OrderedDict((g[0], [0, 0, 0]) for g in groups)
        other = [0, 0, 0]
        for l in lines:
            for nm, f, lo, hi in groups:
                if l[0].startswith(f) and lo <= l[1] <= hi:
                    a = acc[nm]; break
            else:
                a = other
            a[0] += l[3]; a[1] += l[4]; a[2] += l[5]
        print("== groups for", fn[:80])
        for nm, a in list(acc.items()) + [("other", other)]:
            print("  %-22s %5.1f%% exec  %5.1f thr" % (nm, 100 * a[0] / totE, a[2] / max(a[0], 1)))
        break
