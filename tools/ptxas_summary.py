"""Digest of `nvcc -Xptxas -v` output: registers / spills / stack per kernel instantiation."""
import re
import subprocess
import sys

log = open(sys.argv[1]).read()
ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:ptxas info\s*:\s*Function properties.*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n)?ptxas info\s*:\s*Used (\d+) registers", log)
for name, stack, ss, sl, regs in ents:
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r'void dlpd_kernel<(.*)>\(.*', r'\1', dem)
    print(f"{dem:40s} regs={regs:>3s} stack={stack} spill={ss}/{sl}")
