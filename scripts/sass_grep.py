"""profiles/rNN_sass_grep.txt: instruction-mnemonic counts per kernel of the shipped library (``cuobjdump -sass``).
    python scripts/sass_grep.py > profiles/r02_sass_grep.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "maxsquareloss_b200", "lib", "libmsq_b200.so")
KEYS = ["FFMA2", "FMUL2", "FADD2", "MUFU.EX2", "MUFU.LG2", "MUFU.RCP", "SHF", "LDGSTS", "LDG.E.128", "LDG.E.64", "LDS.128", "LDS.64",
        "STS.64", "ATOMS", "ATOMG", "RED.E", "REDUX", "UTMALDG", "UBLKCP", "UTCMMA", "SYNCS", "ACQBULK", "BAR.SYNC", "SHFL", "LDL", "STL"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
per, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k in ("LDG.E.128", "LDG.E.64", "LDS.128", "LDS.64", "STS.64", "MUFU.EX2", "MUFU.LG2", "MUFU.RCP", "RED.E", "BAR.SYNC") and op.startswith(k)):
                per[cur][k] += 1
                break
names = subprocess.run(["cu++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print(f"# SASS of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; architectures in the file: {', '.join(archs)}), instruction mnemonic counts")
print("# packed fp32 (FFMA2/FMUL2/FADD2) is Blackwell-only; LDGSTS = cp.async; SHF = funnel shifts (the near-maximum class word);")
print("# UTMALDG/UBLKCP (TMA) and UTCMMA (tcgen05) are deliberately absent: nothing on this path is a contraction, and the staged")
print("# tile is <= 8 KB with a 516-byte row pitch (TMA tensor maps need 16-byte pitches); LDL/STL = local memory (cold tie-replay")
print("# arrays, a few spills)\n")
print("total: " + "  ".join(f"{k} {tot[k]}" for k in KEYS) + "\n")
for (mangled, c), name in zip(per.items(), names):
    print(name[:110])
    print("    " + "  ".join(f"{k} {c[k]}" for k in KEYS if c[k]))
