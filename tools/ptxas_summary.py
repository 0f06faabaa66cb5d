#!/usr/bin/env python3
"""Summarise `-Xptxas -v` logs written by csrc/Makefile: registers, stack, spills per kernel."""
import pathlib, re, subprocess, sys
root = pathlib.Path(__file__).resolve().parents[1] / "optix_raytracer_b200" / "csrc" / "build"
for f in sorted(root.glob("*.ptxas.log")):
    txt = f.read_text()
    for w in re.findall(r".*(?:warning|error).*", txt)[:5]:
        print("  !!", w)
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        print(f"{f.stem.split('.')[0]:12s} {name[:64]:64s} regs={m.group(5):>3s} stack={m.group(2):>4s} spill={m.group(3)}/{m.group(4)}")
