"""Static SASS statistics of one kernel: total instructions, the innermost-to-outermost loops (backward
branches) with their opcode histograms.  Usage: python scripts/sass_count.py lib.so kernel_substring [--dump]"""
import collections
import re
import subprocess
import sys


def kernel_sass(lib, name):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    blocks = out.split("Function : ")
    hits = [b for b in blocks[1:] if name in b.split("\n", 1)[0]]
    if not hits:
        raise SystemExit(f"no kernel matching {name!r}")
    return hits[0]


def parse(block):
    ins = []
    for line in block.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def opcode(text):
    t = re.sub(r"^@!?U?P\d+\s+", "", text)
    return t.split()[0].split(".")[0]


def main():
    lib, name = sys.argv[1], sys.argv[2]
    block = kernel_sass(lib, name)
    print(block.split("\n", 1)[0])
    ins = parse(block)
    print("total instructions:", len(ins))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    for lo, hi in sorted(loops, key=lambda x: x[1] - x[0]):
        body = ins[lo:hi + 1]
        hist = collections.Counter(opcode(t) for _, t in body)
        print(f"loop 0x{ins[lo][0]:x}..0x{ins[hi][0]:x}: {len(body)} instructions")
        print("   " + "  ".join(f"{k}:{v}" for k, v in hist.most_common(28)))
    if "--dump" in sys.argv:
        for a, t in ins:
            print(f"{a:05x}  {t}")


if __name__ == "__main__":
    main()
