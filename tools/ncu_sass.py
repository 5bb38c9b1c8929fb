"""SASS listing of one kernel from an ncu report, in program order, with executed-instruction counts, stall
samples and the source line of every instruction.  usage: python tools/ncu_sass.py report.ncu-rep kernel_sub [cubin_sub]"""
import csv, io, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines
rep, ksub = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[h]; ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
inst = [(int(r[ci]), int(r[cs]), r[1]) for r in rows[h + 1:] if len(r) > ci and r[ci].isdigit()]
sass = ncu_lines.sass_lines(ksub, sys.argv[3] if len(sys.argv) > 3 else "traverse_warp")
for (n, smp, txt), (addr, key, _) in zip(inst, sass):
    print("%6x %10d %6d  %-22s %s" % (addr, n, smp, ("%s:%d" % (key[0][:14], key[1])) if key else "?", txt[:110]))
