#!/usr/bin/env python
"""SASS opcode summary of the library's kernels (cuobjdump -sass of the in-tree objects): per kernel family the number of
instantiations and, for one representative each, total instructions and the counts of the opcodes that prove the design
(UBLKCP / SYNCS: cp.async.bulk through the TMA unit + mbarriers; LDGSTS: cp.async; UCGABAR / MAPA-style cluster ops; DFMA
outside division / sqrt sequences would betray an FMA contraction).   python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections, glob, os, re, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UBLKCP", "SYNCS", "LDGSTS", "LDGDEPBAR", "BAR", "UCGABAR_ARV", "UCGABAR_WAIT", "LDS", "STS", "LDG", "STG", "CCTL", "DADD", "DMUL",
       "DFMA", "MUFU", "IMAD", "RED", "ATOM", "ATOMG", "WARPSYNC", "ELECT"]
rows = []
for obj in sorted(glob.glob(os.path.join(ROOT, "swalbe.jl_b200", "csrc", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    names = subprocess.run(["c++filt"], input="\n".join(f.split("\n", 1)[0].strip() for f in funcs), capture_output=True, text=True).stdout.split("\n")
    for f, name in zip(funcs, names):
        name = re.sub(r"\(.*", "", name.replace("swalbe::", "").replace("(anonymous namespace)::", "").replace("void ", ""))
        ops = [re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in re.findall(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", f)]
        rows.append((os.path.basename(obj), name, len(ops), collections.Counter(ops)))
print("SASS opcode summary (sm_100a cubins of the in-tree objects; static counts per kernel)\n")
hdr = f"{'kernel':84s} {'instr':>6s} " + " ".join(f"{o[:7]:>7s}" for o in OPS)
print(hdr)
seen = set()
for obj, name, n, c in sorted(rows, key=lambda r: (r[1], r[0])):
    fam = re.sub(r"<.*", "", name)
    key = name
    if fam == "k_fused_step":  # one line per flavour at one width (224 where it exists)
        m = re.match(r"k_fused_step<(\d+), \d+, (.*)>", name)
        if not m or m.group(1) != "224":
            continue
        key = "k_fused_step<224, *, " + m.group(2) + ">"
        if not re.search(r", (1|-1), ", ", " + m.group(2) + ", "):  # pressure mode BROAD_93 or the full kernel only
            continue
    if key in seen:
        continue
    seen.add(key)
    print(f"{key[:84]:84s} {n:6d} " + " ".join(f"{c.get(o, 0):7d}" for o in OPS))
fams = collections.Counter(re.sub(r"<.*", "", r[1]) for r in rows)
print("\ninstantiations per kernel family: " + ", ".join(f"{k} x{v}" for k, v in sorted(fams.items())))
print("template parameters of k_fused_step: <NT, min CTAs/SM, tau==1, thermal, pressure mode (-1 = run-time options, FULL), "
      "BULK (TMA rows), g==0, OPTS, FM (tau != 1 from moments), NS (neighbour sync)>")
