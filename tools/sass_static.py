"""Static SASS opcode histogram per kernel of libb200vmaf.so (cuobjdump -sass): the opcode evidence for profiles/.
Usage: python tools/sass_static.py [lib.so] > profiles/r02_sass_static.md"""
import collections, os, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pqa2_b200", "libb200vmaf.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, fn = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        kern[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and fn:
        op = m.group(1)
        if op in ("FFMA2", "FMUL2", "FADD2", "REDUX", "IDP", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "I2F", "F2I"):
            pass
        kern[fn][op] += 1
names = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
def short(n):
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    n = re.sub(r"\(BvBatch.*", "", n)
    return n.replace("void ", "")
total = collections.Counter()
for c in kern.values():
    total.update(c)
print("# Static SASS opcode counts per kernel (cuobjdump -sass pqa2_b200/libb200vmaf.so)\n")
print(f"cubin architectures: {', '.join(archs)}; {len(kern)} kernels; {sum(total.values())} SASS instructions.\n")
key = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "DFMA", "DMUL", "DADD", "IMAD", "IADD3", "IDP", "REDUX", "MUFU", "LDS", "STS", "LDG", "STG", "BAR"]
print("whole library: " + ", ".join(f"`{k}` {total[k]}" for k in key if total[k]) + "\n")
absent = [k for k in ("UTMALDG", "UTMASTG", "UTCMMA", "UTCHMMA", "LDTM", "STTM", "HMMA", "IMMA", "UBLKCP") if not any(o.startswith(k) for o in total)]
print("absent (no tensor-core / TMEM / TMA instruction anywhere, as the north star prescribes for these stencils): " + ", ".join(absent) + "\n")
print("kernel | instr | " + " | ".join(key))
print("--- | --- | " + " | ".join("---" for _ in key))
for n, c in zip(names, kern.values()):
    print(f"{short(n)} | {sum(c.values())} | " + " | ".join(str(c[k]) if c[k] else "" for k in key))
