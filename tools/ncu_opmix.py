"""Instruction mix and stall-sample share by SASS opcode for one kernel of an .ncu-rep (needs --import-source on / set full)."""
import csv, collections, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
ops = collections.defaultdict(lambda: [0, 0])
tot_s = tot_i = 0
for r in rows:
    if "Source" in r and "# Samples" in r:
        if hdr is not None:
            break                      # first kernel instance only
        hdr = r
        iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= iI:
        continue
    parts = r[iS].strip().split()
    if not parts:
        continue
    op = parts[0] if not parts[0].startswith('@') else parts[1]
    op = op.split('.')[0]
    try:
        s, n = int(r[iN] or 0), int(r[iI] or 0)
    except ValueError:
        continue
    ops[op][0] += s; ops[op][1] += n; tot_s += s; tot_i += n
print("total samples", tot_s, "warp instr", tot_i)
for op, (s, n) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"{op:10s} samples {100*s/max(tot_s,1):5.1f}%  instr {100*n/max(tot_i,1):5.1f}%")
