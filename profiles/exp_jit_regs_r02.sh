timeout 300 env MS_JIT_MAXREG=0 python profiles/bench_brainfuck.py --burner 40 40 60 2>&1 | tail -1
timeout 300 env MS_JIT_MAXREG=64 python profiles/bench_brainfuck.py --burner 40 40 60 2>&1 | tail -1
timeout 300 env MS_JIT_MAXREG=96 python profiles/bench_brainfuck.py --burner 40 40 60 2>&1 | tail -1
timeout 300 env MS_JIT_MAXREG=128 python profiles/bench_brainfuck.py --burner 40 40 60 2>&1 | tail -1
timeout 300 env MS_JIT_MAXREG=168 python profiles/bench_brainfuck.py --burner 40 40 60 2>&1 | tail -1
