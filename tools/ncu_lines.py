"""Per-source-line instruction and stall-sample totals of one kernel in an .ncu-rep (needs -lineinfo and --import-source on)."""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, done = None, None, set()
acc = collections.defaultdict(lambda: [0, 0, ""])
tot_i = tot_s = 0
nfunc = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name":
        continue
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; iL, iS = 0, 1; iA = 2; iN, iI = hdr.index("# Samples"), hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) <= iI or r[iA] != "-":
        continue
    try:
        n, i = int(r[iN] or 0), int(r[iI] or 0)
    except ValueError:
        continue
    key = (cur_file, int(r[iL]))
    acc[key][0] += n; acc[key][1] += i; acc[key][2] = r[iS].strip()[:110]
    tot_s += n; tot_i += i
print("total samples", tot_s, "warp instr", tot_i)
for (f, ln), (n, i, src) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*i/max(tot_i,1):5.1f}% instr {100*n/max(tot_s,1):5.1f}% smp  {f}:{ln}  {src}")
