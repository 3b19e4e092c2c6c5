"""Hot source lines of one kernel in an ncu report (needs -lineinfo + --import-source on):
   python scripts/ncu_lines.py report.ncu-rep kernel-regex [samples|inst] [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
key = sys.argv[3] if len(sys.argv) > 3 else "samples"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
cur = None; agg = []
for r in csv.reader(out.splitlines()):
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 7 and r[0].isdigit():
        try: agg.append((int(r[6]) if r[6].isdigit() else 0, int(r[7]), cur, int(r[0]), r[1].strip()[:100]))
        except ValueError: pass
ts = sum(a[0] for a in agg); ti = sum(a[1] for a in agg)
print('samples', ts, 'instructions', ti)
agg.sort(key=lambda a: a[0] if key == "samples" else a[1], reverse=True)
for a in agg[:top]:
    print('%6d %5.1f%% inst %9d %5.1f%%  %s:%d  %s' % (a[0], 100 * a[0] / max(ts, 1), a[1], 100 * a[1] / max(ti, 1), a[2], a[3], a[4]))
