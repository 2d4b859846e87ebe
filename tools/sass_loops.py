#!/usr/bin/env python
"""Static view of the hot loops of a kernel in a cubin / shared library (no GPU needed): for every
loop (backward branch) with enough FP64 work, the instructions on its hot path, their mix, and the sum
of the stall counts the scheduler encoded into them (bits 41-44 of the second instruction word) --
the least number of cycles ONE warp needs for an iteration, whatever the other warps do. Basic blocks
that contain a CALL (the IEEE division's slow path, out-of-line helpers) count as cold.

    python tools/sass_loops.py pyfds_b200/libfdsb200.so streamv_kernelILi2ELb1ELb1ELi2 [--dump ADDR]
    python tools/sass_loops.py pyfds_b200/libfdsb200.so streamv --ranges DUMPFILE lo-hi [lo-hi ...]

The first form lists the loops of every kernel whose mangled name contains the second argument;
--dump ADDR prints the loop that starts at ADDR block by block (cold blocks marked C). The second form
sums over address ranges of such a dump (to separate the fast body of a loop from the exact one).
DESIGN.md 4.2b and profiles/README.md quote numbers of this tool."""

import collections
import re
import subprocess
import sys

INSTR = re.compile(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* 0x([0-9a-f]+) \*/')
WORD = re.compile(r'/\* 0x([0-9a-f]+) \*/')
FP64 = ('DADD', 'DMUL', 'DFMA')


def opcode(text):
    text = re.sub(r'^@!?U?P\d+\s+', '', text)
    return text.split()[0].split('.')[0]


def functions(path):
    sass = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    for chunk in re.split(r'\n\s+Function : ', sass)[1:]:
        name, lines = chunk.split('\n', 1)[0], chunk.split('\n')
        code = []
        for k, line in enumerate(lines):
            m = INSTR.match(line)
            if not m:
                continue
            control = int(WORD.search(lines[k + 1]).group(1), 16)
            code.append((int(m.group(1), 16), m.group(2), (control >> 41) & 0xF))
        yield name, code


def summary(code):
    mix = collections.Counter(opcode(text) for _, text, _ in code)
    moves = sum(1 for _, text, _ in code if 'IMAD.MOV' in text or re.match(r'(@\S+\s+)?MOV ', text))
    stall = sum(s for _, _, s in code)
    return ('n=%d stall=%d (%.2f per instruction) fp64=%d mov=%d shfl=%d lds=%d ldl=%d stl=%d' % (
        len(code), stall, stall / max(len(code), 1), sum(mix[o] for o in FP64), moves, mix['SHFL'],
        mix['LDS'], mix['LDL'], mix['STL']))


def loops(code, min_fp64=150, max_len=2600):
    index = {a: k for k, (a, _, _) in enumerate(code)}
    targets = set()
    for _, text, _ in code:
        if re.search(r'\b(BRA|BSSY|CALL|JMP)', text):
            targets.update(int(m, 16) for m in re.findall(r'0x([0-9a-f]+)', text)
                           if int(m, 16) in index)
    for k, (addr, text, _) in enumerate(code):
        m = re.search(r'\bBRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)\)?', text)
        if not m or int(m.group(1), 16) > addr or int(m.group(1), 16) not in index:
            continue
        body = code[index[int(m.group(1), 16)]:k + 1]
        if sum(1 for _, t, _ in body if opcode(t) in FP64) < min_fp64 or len(body) > max_len:
            continue
        blocks, block = [], []
        for item in body:
            if item[0] in targets and block:
                blocks.append(block)
                block = []
            block.append(item)
            if re.search(r'\b(BRA|CALL|RET|EXIT|BSYNC)\b', item[1]):
                blocks.append(block)
                block = []
        if block:
            blocks.append(block)
        yield int(m.group(1), 16), addr, blocks


def main():
    path, needle = sys.argv[1], sys.argv[2]
    if '--ranges' in sys.argv:
        k = sys.argv.index('--ranges')
        ranges = [tuple(int(x, 16) for x in r.split('-')) for r in sys.argv[k + 2:]]
        code = []
        for line in open(sys.argv[k + 1]):
            m = re.match(r'(C?)\s+([0-9a-f]+) st=(\d+) (.*)', line)
            if m and not m.group(1) and any(lo <= int(m.group(2), 16) <= hi for lo, hi in ranges):
                code.append((int(m.group(2), 16), m.group(4), int(m.group(3))))
        print(summary(code))
        return
    dump = int(sys.argv[sys.argv.index('--dump') + 1], 16) if '--dump' in sys.argv else None
    for name, code in functions(path):
        if needle not in name:
            continue
        print(name, len(code), 'instructions')
        for start, end, blocks in loops(code):
            cold = [any('CALL' in text for _, text, _ in b) for b in blocks]
            hot = [item for b, c in zip(blocks, cold) if not c for item in b]
            print('  loop %x-%x: %d instructions, hot path: %s' % (
                start, end, sum(len(b) for b in blocks), summary(hot)))
            if dump == start:
                for b, c in zip(blocks, cold):
                    for addr, text, stall in b:
                        print('%s %x st=%d %s' % ('C' if c else ' ', addr, stall, text))
                    print('  ----')


if __name__ == '__main__':
    main()
