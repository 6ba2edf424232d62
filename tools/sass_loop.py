#!/usr/bin/env python
"""Prints the SASS of one kernel of libslamgpu.so with the scheduling control fields decoded (stall count, write / read
scoreboard, wait mask) -- the fields cuobjdump only shows as hex.  Usage: tools/sass_loop.py <kernel-name-substring> [lo hi]"""
import os, re, subprocess, sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "slam_constructor_b200", "lib", "libslamgpu.so")


def decode(name):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.split("\n")
    out, on, i = [], False, 0
    while i < len(txt):
        ln = txt[i]
        if "Function :" in ln:
            on = name in ln
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", ln)
            if m and i + 1 < len(txt):
                m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", txt[i + 1])
                if m2:
                    hi = int(m2.group(1), 16)
                    out.append((m.group(1), m.group(2).strip(), (hi >> 41) & 0xF, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3F))
                    i += 2
                    continue
        i += 1
    return out


if __name__ == "__main__":
    rows = decode(sys.argv[1])
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
    for k, o in enumerate(rows[lo:hi], lo):
        print("%4d %s %-62s stall %2d wr %s rd %s wait %s" % (k, o[0], o[1][:62], o[2], o[3] if o[3] != 7 else "-", o[4] if o[4] != 7 else "-",
                                                                format(o[5], "06b")))
