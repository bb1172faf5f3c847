#!/usr/bin/env python
"""Static SASS statistics of one kernel instantiation: total instructions, the longest backward-branch loop body (the
steady-state row iteration of k_fused_step) and its opcode histogram.
   python tools/sass_loop_count.py swalbe.jl_b200/csrc/fused_v224.o 'k_fused_step<224, 3, true, false, 1, true, true, false, false>'"""
import collections, re, subprocess, sys

obj, want = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)[1:]
names = subprocess.run(["c++filt"], input="\n".join(f.split("\n", 1)[0].strip() for f in funcs), capture_output=True, text=True).stdout.split("\n")
for f, name in zip(funcs, names):
    name = name.replace("swalbe::", "").replace("(swalbe::FusedArgs)", "").replace("void ", "")
    if want not in name:
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", f)
    addr = [int(a, 16) for a, _ in ins]
    ops = [t for _, t in ins]
    best = (0, 0, 0)
    for k, t in enumerate(ops):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr[k]:
                j = addr.index(tgt) if tgt in addr else None
                if j is not None and k - j > best[0]:
                    best = (k - j + 1, j, k)
    n, j, k = best
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in ops[j:k + 1])
    print(f"{name}\n  total {len(ops)} instructions; longest loop body {n} instructions")
    print("  " + ", ".join(f"{o} {c}" for o, c in hist.most_common(28)))
