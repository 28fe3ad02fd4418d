"""SASS evidence for the tcgen05 / TMA / PDL claims: per kernel of libsdt_b200.so, how often the Blackwell-specific instructions
occur (``cuobjdump -sass``; runs without a GPU).  `python tools/sass_histogram.py > profiles/r02_sass_histogram.txt`

    UTCHMMA[.2CTA]   tcgen05.mma kind::f16 (bf16 / fp16 operands, f32 accumulate in TMEM); .2CTA = cta_group::2
    UTMALDG / UTMASTG cp.async.bulk.tensor (TMA tile loads / stores); UTMACCTL = prefetch.tensormap
    UTCCP            tcgen05.cp (shared -> tensor memory; the opt-in TS-mode main loop)
    LDTM             tcgen05.ld (TMEM -> registers)            UTCBAR  tcgen05.commit -> mbarrier
    SYNCS            mbarrier arrive / try_wait               UTCATOMSWS  tcgen05.alloc / dealloc
    ACQBULK / PREEXIT griddepcontrol.wait / .launch_dependents (programmatic dependent launch)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "scal_sdt_b200", "_build", "libsdt_b200.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCCP", "UTMALDG", "UTMASTG", "UTMACCTL", "LDTM", "UTCBAR", "SYNCS", "UTCATOMSWS", "ACQBULK", "PREEXIT", "RED", "ATOM",
        "HMMA", "FFMA", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        c = kernels[cur]
        c["_total"] += 1
        if op.startswith("UTCHMMA"):
            c["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
        else:
            base = op.split(".")[0]
            if base in KEYS:
                c[base] += 1
            elif base.startswith("ATOM"):
                c["ATOM"] += 1
            elif base.startswith("RED"):
                c["RED"] += 1
            elif base.startswith("UTMACCTL") or base.startswith("UBLKPF") or "PREFETCH" in base:
                c["UTMACCTL"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  ({len(kernels)} kernels, sm_100a)")
    print(f"{'instr':>7} " + " ".join(f"{k.replace('UTCHMMA.2CTA', 'MMA.2CTA').replace('UTCATOMSWS', 'TMEMALLOC'):>9}" for k in KEYS) + "  kernel")
    tot = collections.Counter()
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("void ", "").replace("sdt::", "")
        print(f"{c['_total']:>7} " + " ".join(f"{c[k]:>9}" for k in KEYS) + f"  {short}")
        tot.update(c)
    print(f"{tot['_total']:>7} " + " ".join(f"{tot[k]:>9}" for k in KEYS) + "  TOTAL")


if __name__ == "__main__":
    sys.exit(main())
