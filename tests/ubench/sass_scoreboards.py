#!/usr/bin/env python
"""Which scoreboard does ptxas give each long-latency instruction of a kernel, and who waits for it?

    python tests/ubench/sass_scoreboards.py graspbalance_b200/csrc/scatter_private.o ILi2ELi32ELi4ELi2ELi6ELi512

Decodes the control bits of sm_100a SASS as printed by `cuobjdump -sass` (two 64-bit words per instruction; bits 41-63 of
the second word: stall count [41:44], yield [45], write barrier [46:48], read barrier [49:51], wait mask [52:57]) and lists
every global load with the scoreboard it sets, plus every instruction that waits for one of those scoreboards.  This is how
the single-scoreboard drain of the group backward's register ring was found."""
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for f in txt.split("Function : ")[1:]:
        name = f.split("\n", 1)[0]
        if pat not in name:
            continue
        lines, ins, i = f.split("\n"), [], 0
        while i < len(lines):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
            if m and i + 1 < len(lines):
                hi = int(re.search(r"/\* (0x[0-9a-f]+) \*/", lines[i + 1]).group(1), 16)
                c = hi >> 41
                ins.append((int(m.group(1), 16), m.group(2).strip(), c & 15, (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 63))
                i += 2
            else:
                i += 1
        print(name, len(ins), "instructions")
        ldg_sb = sorted({w for a, s, st, w, r, wait in ins if s.lstrip("@!P0123456 ").startswith("LDG") and w != 7})
        print("scoreboards set by global loads:", ldg_sb)
        for a, s, st, w, r, wait in ins:
            if "LDG" in s or "CALL" in s or any(wait >> b & 1 for b in ldg_sb):
                print(f"{a:6x}  stall {st:2d}  sets {'-' if w == 7 else w}  waits {wait:06b}  {s[:90]}")


if __name__ == "__main__":
    main()
