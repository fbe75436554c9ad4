"""profiles/r2_sass_excerpts.txt: per hot kernel the SASS size, mnemonic histogram and the first instructions of the kinds that
matter (fp64 FMA / DMMA, reductions, asynchronous copies, vector loads, shuffles, barriers). Usage: python scripts/sass_excerpts.py"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "l3ster_b200", "csrc", "build")
report = []


def excerpt(obj, pattern, title, mnems, max_lines=6):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
    funcs = [m.start() for m in re.finditer(r"Function : ", txt)]
    for i, st in enumerate(funcs):
        en = funcs[i + 1] if i + 1 < len(funcs) else len(txt)
        head = txt[st:st + 800].split("\n")[0]
        if not re.search(pattern, head):
            continue
        lines = [ln for ln in txt[st:en].split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/", ln)]
        report.append(f"## {title}\n{head[:170]}\n{len(lines)} SASS instructions ({16 * len(lines) / 1024:.0f} kB)")
        counts = {}
        for ln in lines:
            m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
            if m:
                op = m.group(1).split(".")[0]
                counts[op] = counts.get(op, 0) + 1
        report.append("mnemonic counts: " + ", ".join(f"{k} {v}" for k, v in sorted(counts.items(), key=lambda kv: -kv[1])[:16]))
        for mn in mnems:
            sel = [ln.strip() for ln in lines if re.search(mn, ln)]
            report.append(f"-- /{mn}/: {len(sel)} instructions" + (", first:" if sel else ""))
            report.extend("    " + ln[:150] for ln in sel[:max_lines])
        report.append("")
        return
    report.append(f"## {title}: not found in {obj}\n")


excerpt("kernels/reg_diffusion3d.o", r"mfHexPlanesKernel.*Li4ELi5ELi1ELb0E", "mfHexPlanesKernel<bench_diffusion3d, P=4, NQ=5> — matrix-free apply",
        [r"DFMA", r"REDG", r"LDGSTS", r"LDG\.E.*(256|ENL2)", r"SHFL", r"BAR\."])
excerpt("kernels/reg_diffusion3d.o", r"assembleDmmaKernel.*Li3ELi4E", "assembleDmmaKernel<bench_diffusion3d, hex P=4> — assembly + CRS scatter",
        [r"DMMA\.", r"REDG", r"LDS", r"BAR\."])
excerpt("capi.o", r"condenseKernelILi4E", "condenseKernel<4> — static condensation", [r"DFMA", r"REDG", r"LDS\.(128|64)"], 4)
excerpt("capi.o", r"condenseLargeKernel", "condenseLargeKernel — static condensation, more than 256 interior dofs", [r"DFMA", r"REDG"], 4)
excerpt("capi.o", r"cgUpdateKernel", "cgUpdateKernel — CG vector pass", [r"LDG\.E\.(128|64)", r"STG", r"DFMA"], 4)
excerpt("capi.o", r"haloPackKernel", "haloPackKernel — Import packing", [r"LDG", r"STG"], 3)
out = os.path.join(ROOT, "profiles", "r2_sass_excerpts.txt")
open(out, "w").write("SASS excerpts of the hot kernels (cuobjdump -sass of the objects under l3ster_b200/csrc/build; sm_100a, nvcc 12.9).\n"
                     "All arithmetic of this path is fp64: there is no tcgen05 (UTCMMA) form of it; DMMA (mma.sync.m8n8k4.f64) is the tensor-core\n"
                     "instruction for the B^T W B contraction, DFMA for everything else. LDGSTS = cp.async staging, RED = fire-and-forget atomics.\n\n"
                     + "\n".join(report))
print(open(out).read()[:5000])
