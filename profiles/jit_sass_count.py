#!/usr/bin/env python
"""Static SASS figures of the specialised evaluator kernels, no GPU needed: the source `ms_eval_jit_check` generates for a
program (MS_EVAL_JIT_DUMP) is compiled to a cubin with NVRTC for sm_100a (cuda-python) and disassembled with cuobjdump;
the kernels are straight-line code (one pass, no loops except the Fermat / power chains, which are unrolled calls), so the
instruction count is the per-thread issue count.

    python profiles/jit_sass_count.py [repo_root ...]      # default: this checkout; give an older checkout to compare

Prints one markdown table row per program: instructions, the multiply pipe (IMAD*), registers."""
import collections
import ctypes
import os
import re
import subprocess
import sys
import tempfile

from cuda.bindings import nvrtc


def compile_cubin(src):
    err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"ms_eval_jit.cu", 0, [], [])
    assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS
    opts = [b"--gpu-architecture=sm_100a", b"-std=c++17"]
    err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if err != nvrtc.nvrtcResult.NVRTC_SUCCESS:
        _, n = nvrtc.nvrtcGetProgramLogSize(prog)
        log = b" " * n
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise RuntimeError(log.decode()[:2000])
    _, n = nvrtc.nvrtcGetCUBINSize(prog)
    cubin = b" " * n
    nvrtc.nvrtcGetCUBIN(prog, cubin)
    return cubin


def sass_stats(cubin):
    with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
        f.write(cubin)
        path = f.name
    try:
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    finally:
        os.unlink(path)
    ops = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1).split(".")[0]] += 1
    regs = re.findall(r"REG:(\d+)", res)
    return ops, (int(regs[0]) if regs else None)


def main(root):
    sys.path.insert(0, root)
    for m in [k for k in sys.modules if k.startswith("ministark_b200")]:
        del sys.modules[m]
    from ministark_b200 import _lib
    from ministark_b200.air import Air, ProofOptions
    from ministark_b200.examples import brainfuck as bf
    from ministark_b200.examples import fib
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    _lib.LIB_PATH = os.path.join(here, "ministark_b200", "libministark_b200.so")     # the generator of THIS checkout's library
    lib = _lib.load()
    trace, out = bf.simulate(bf.HELLO_WORLD)
    claim = bf.BrainfuckClaim(bf.HELLO_WORLD, b"", out)
    air = Air(claim.AirConfig, 1 << 20, claim, ProofOptions(19, 16, 20, 16, 16))
    fa = Air(fib.FibAirConfig, 1 << 21, 5, fib.OPTIONS)
    cases = [("brainfuck composition", air.composition_program(), 3), ("brainfuck DEEP", air.deep_program()[0], 3),
             ("fib composition", fa.composition_program(), 1), ("fib DEEP", fa.deep_program()[0], 1)]
    for name, sym, fq in cases:
        p = sym.bind(challenges=[(1, 2, 3)] * 32, hints=[(4, 5, 6)] * 256, ccoefs=[(7, 8, 9)] * 256)
        with tempfile.NamedTemporaryFile(suffix=".cu", delete=False) as f:
            path = f.name
        os.environ["MS_EVAL_JIT_DUMP"] = path
        log = ctypes.create_string_buffer(1 << 14)
        assert lib.ms_eval_jit_check(p.code.ctypes.data, len(p), p.consts.ctypes.data, p.consts.shape[0], fq, log, 1 << 14) == 0, log.value
        os.environ.pop("MS_EVAL_JIT_DUMP")
        src = open(path).read()
        os.unlink(path)
        ops, regs = sass_stats(compile_cubin(src))
        total = sum(ops.values())
        imad = sum(v for k, v in ops.items() if k.startswith("IMAD") or k.startswith("IMUL"))
        alu = sum(v for k, v in ops.items() if k in ("IADD3", "IADD", "LOP3", "SHF", "SEL", "ISETP", "LEA", "PRMT", "MOV", "IADD32I", "UIADD3"))
        mem = sum(v for k, v in ops.items() if k in ("LDG", "STG", "LDC", "LD", "ST", "LDL", "STL"))
        print(f"| {os.path.basename(os.path.abspath(root))} | {name} | {len(p)} | {total} | {imad} | {alu} | {mem} | {ops.get('CALL', 0) + ops.get('BRA', 0)} | {regs} |")


if __name__ == "__main__":
    print("| checkout | program | program instructions | SASS instructions | IMAD/IMUL | integer ALU | memory | calls + branches | registers |")
    print("|---|---|---|---|---|---|---|---|---|")
    for r in (sys.argv[1:] or [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]):
        main(r)
