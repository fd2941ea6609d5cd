"""Print a timeline CSV of profiles/timeline.py as text: start, duration, end, stream, kernel.   python profiles/show_timeline.py file.csv [from_us]"""
import csv, re, sys
rows = list(csv.DictReader(open(sys.argv[1])))
t0 = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
for r in rows:
    s, d = float(r['start_us']), float(r['dur_us'])
    if s < t0: continue
    n = re.sub(r'\(.*', '', r['name']).replace('void acvae::', '').replace('acvae::', '')[:48]
    print(f"{s:8.1f} {d:7.1f} {s + d:8.1f} s{r['stream']:>3} {n}")
