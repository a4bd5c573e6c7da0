"""python tools/sass_histogram.py [out.txt]: opcode histogram of every kernel in primerforge_b200/libpfb200.so (cuobjdump -sass);
the mnemonics that matter on sm_100a are called out: UBLKCP (cp.async.bulk), SYNCS (mbarrier), LDGSTS (cp.async), REDG / ATOMG,
no UTC*MMA / HMMA (nothing here is a contraction)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "primerforge_b200", "libpfb200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        full = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = full.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        hist.setdefault(kern, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
out = []
total = collections.Counter()
for k, c in hist.items():
    total.update(c)
    n = sum(c.values())
    out.append(f"{k}: {n} instructions: " + " ".join(f"{op} {v}" for op, v in c.most_common(12)))
special = ("UBLKCP", "SYNCS", "LDGSTS", "REDG", "ATOMG", "RED", "ATOM", "UTMALDG", "HMMA", "UTCHMMA", "VOTE", "POPC", "SHF", "IMAD", "LOP3", "ISETP", "LDS", "STS", "LDG")
out.append("")
out.append("whole library: " + " ".join(f"{op} {total.get(op, 0)}" for op in special))
text = "\n".join(out) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
print(text[-1500:])
