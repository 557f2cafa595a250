"""Top source lines by warp-stall samples from an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py <report.ncu-rep> [kernel-substring] [top-n]"""
import csv, subprocess, sys
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur_file = cur_fn = None; hdr = None; out = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < len(hdr) or (want and want not in (cur_fn or "")): continue
    if not r[0].strip().isdigit(): continue      # SASS rows have an empty line number
    try: n = int(r[hdr["# Samples"]])
    except ValueError: continue
    if n == 0: continue
    st = {k[6:]: int(r[i]) for k, i in hdr.items() if k.startswith("stall_") and "Not Issued" not in k and r[i].isdigit() and int(r[i]) > 0}
    out.setdefault(cur_fn, []).append((n, cur_file.split("/")[-1], r[0], r[1].strip()[:110], int(r[hdr["Instructions Executed"]]), st))
for fn, L in out.items():
    tot = sum(x[0] for x in L)
    print(f"== {fn[:100]}  total samples {tot}")
    for n, f, ln, txt, ie, st in sorted(L, key=lambda x: -x[0])[:topn]:
        top = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{100*n/tot:5.1f}%  {f}:{ln:>4}  inst={ie:>9}  [{top}]  {txt}")
