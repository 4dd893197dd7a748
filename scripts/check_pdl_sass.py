"""Disassemble libtaxo_sm100.so and check every kernel that waits on its predecessor (ACQBULK = griddepcontrol.wait): no instruction that
touches global memory may be scheduled in front of the wait (the compiler hoists non-coherent loads over inline asm, see tx_common.cuh).
usage: python scripts/check_pdl_sass.py [path/to/lib.so]   -> prints one line per kernel, exit code 1 on a violation"""
import re
import subprocess
import sys

# global-memory instructions.  Not counted: generic volatile loads (LD.E.STRONG.SYS: the GEMM prologue reads its TMEM slot from shared
# memory through a generic pointer), mbarrier traffic (SYNCS.*) and the CCTL.IVALL of the prologue's mbarrier-init fence - none of them
# reads data a predecessor kernel produces.
MEM = re.compile(r"^(@!?U?P\d+\s+)?(LDG|LD\.E(?!\.STRONG\.SYS)|STG|ST\.E|ATOMG|ATOM\b|RED\b|REDG|LDGSTS|UBLKCP|UTMALDG|UTMASTG|UBLKPF)")


def check(lib):
    return check_text(subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout)


def check_text(out):
    """{kernel: [global-memory instructions in front of its ACQBULK]} for every kernel of a `cuobjdump -sass` listing that has one"""
    fn, seen_mem, results, waited = None, [], {}, False
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn, seen_mem, waited = m.group(1), [], False
            continue
        if fn is None or waited:
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(2)
        if "ACQBULK" in ins:
            results[fn] = list(seen_mem)
            waited = True
        elif MEM.search(ins):
            seen_mem.append(ins.strip())
    return results


if __name__ == "__main__":
    lib = sys.argv[1] if len(sys.argv) > 1 else "taxoexpan_b200/libtaxo_sm100.so"
    res = check(lib)
    bad = {k: v for k, v in res.items() if v}
    print(f"{len(res)} kernels wait on their predecessor; {len(bad)} with a global-memory instruction in front of the wait")
    for k, v in bad.items():
        print("  VIOLATION", k, v[:3])
    sys.exit(1 if bad else 0)
