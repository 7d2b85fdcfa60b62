import csv, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[0]]
body = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
ci = {n: i for i, n in enumerate(h)}
f = lambda r, n: float(r[ci[n]]) if r[ci[n]] not in ("", "-") else 0.0
tot = sum(f(r, "# Samples") for r in body)
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
# walk through softmax-warp code (ex == 76800 or in softmax region): print cumulative samples between marker instructions
acc = 0.0; accs = {}
start = None
for i, r in enumerate(body):
    s = r[ci['Source']]
    ex = f(r, 'Instructions Executed')
    smp = f(r, '# Samples')
    acc += smp
    for n in stalls:
        accs[n] = accs.get(n, 0) + f(r, n)
    if any(k in s for k in ("LDTM", "STTM", "SYNCS", "VOTE", "BAR", "EXIT", "STG", "UTCBAR", "UTCHMMA", "UTMALDG", "WARPSYNC", "MUFU.RCP")):
        top = sorted(((v, n[6:]) for n, v in accs.items()), reverse=True)[:3]
        print(f"{i:5d} {100*acc/tot:6.2f}%  ex={ex:8.0f}  {s[:60]:60s} " + " ".join(f"{n}:{100*v/tot:.1f}" for v, n in top if v > 0))
        acc = 0.0; accs = {}
