"""Write profiles/r2_sass_excerpts.md: the SASS evidence for the Blackwell-specific instructions of the main kernels (bulk TMA
copies + mbarrier in the fused M^T M kernel, 16-byte volatile peer stores / loads, cp.async prefetch and cluster barriers in the
persistent CG kernels).  Run after a build:  python scripts/sass_excerpts.py"""
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "elphdynamics_b200" / "libelph_b200.so")], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)
want = [("mtm_square_kernelILi1ELi16ELb0ELi256", "mtm_square_kernel<1,16,false,256> (bench kernel: fused M^T M, 32-wide lattice)",
         ["UBLKCP", "SYNCS", "UTMA", "UCGABAR"]),
        ("cgpipe_kernelILi1ELi8ELi128ELi2ELi64ELb0ELb0", "cgpipe_kernel<1,8,128,2,PREFETCH> (pipelined CG, 32x32)",
         ["STG.E.128", "LDG.E.128", "LDGSTS", "UCGABAR", "CGAERRBAR", "BAR.SYNC", "DFMA", "SHFL", "MEMBAR", "LDGDEPBAR", "DEPBAR"]),
        ("cgpipe_kernelILi2ELi4ELi128ELi2ELi64ELb0ELb0", "cgpipe_kernel<2,4,128,2,PREFETCH> (pipelined CG, 64x64 in clusters of 4)",
         ["STG.E.128", "LDG.E.128", "LDGSTS", "UCGABAR", "CGAERRBAR", "BAR.SYNC", "DFMA", "SHFL", "MEMBAR", "LDS", "MAPA", "LD.E"]),
        ("cg_p2p_kernelILi1ELi8ELi256ELb0", "cg_p2p_kernel<1,8,256> (single-reduction CG)", ["STG.E.128", "LDG.E.128", "RED", "MEMBAR", "DFMA"]),
        ("halo_exchange_kernel", "halo_exchange_kernel (peer-memory halo of the sharded products)", ["STG.E.128", "LDG.E.128"])]
out = ["# SASS excerpts (round 2)", "",
       "`cuobjdump -sass elphdynamics_b200/libelph_b200.so`, sm_100a cubins only.  Counts of the instructions that matter per kernel, and the",
       "first occurrences in context.  PTX -> SASS: `cp.async.bulk` = `UBLKCP`, `mbarrier.arrive.expect_tx` = `SYNCS.ARRIVE.TRANS64`,",
       "`st.volatile.global.v2.u64` = `STG.E.128.STRONG.SYS`, `ld.volatile.global.v2.u64` = `LDG.E.128.STRONG.SYS`, `cp.async.cg` = `LDGSTS`,",
       "`barrier.cluster` = `UCGABAR_ARV` / `UCGABAR_WAIT`, DSMEM address mapping = `MAPA` / generic `LD.E` through the shared window.", ""]
for key, title, pats in want:
    body = next((f for f in funcs if f.startswith("_Z") and key in f.split("\n", 1)[0]), None)
    if body is None:
        out += [f"## {title}", "", "(not found in this build)", ""]
        continue
    lines = body.split("\n")
    ops = collections.Counter()
    for ln in lines:
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            ops[m.group(1)] += 1
    out += [f"## {title}", "", f"`{lines[0].strip()}` -- {sum(ops.values())} instructions", "", "| instruction | count |", "|---|---|"]
    shown = set()
    for pat in pats:
        for op, n in sorted(ops.items()):
            if pat in op and op not in shown:
                out.append(f"| `{op}` | {n} |")
                shown.add(op)
    out += ["", "```"]
    seen = set()
    for k, ln in enumerate(lines):
        for pat in pats[:6]:
            if pat in ln and pat not in seen:
                seen.add(pat)
                out += [x.rstrip()[:150] for x in lines[max(0, k - 1):k + 2] if "/*" in x]
                out.append("        ...")
    out += ["```", ""]
(ROOT / "profiles" / "r2_sass_excerpts.md").write_text("\n".join(out) + "\n")
print("\n".join(out)[:3500])
