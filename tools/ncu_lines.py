"""Aggregate an .ncu-rep's per-instruction counters by CUDA source line (needs -lineinfo and --import-source on).
usage: ncu_lines.py <rep> <kernel-regex> [launch index among matching launches] [top N]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + pat,
                      "--launch-skip", str(which), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
stats, fname, hdr, func = {}, "", None, ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        func = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] == "":
        continue
    try:
        i_ie, i_te, i_s = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        ie, te, st = int(r[i_ie]), int(r[i_te]), int(r[i_s] or 0)
    except Exception:
        continue
    s = stats.setdefault((fname, r[0], r[1].strip()[:120]), [0, 0, 0])
    s[0] += ie; s[1] += te; s[2] += st
tot_i = sum(v[0] for v in stats.values()); tot_t = sum(v[1] for v in stats.values()); tot_s = sum(v[2] for v in stats.values())
print(func[:80], "| warp inst", tot_i, "avg threads %.2f" % (tot_t / max(tot_i, 1)), "samples", tot_s)
for key, (ie, te, st) in sorted(stats.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% inst thr %5.2f smp %5.2f%% | %s:%s  %s" % (100 * ie / max(tot_i, 1), te / max(ie, 1), 100 * st / max(tot_s, 1), key[0][:18], key[1], key[2]))
