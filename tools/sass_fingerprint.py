"""Per-kernel SASS fingerprints of a built library, and a diff against a saved set.

Used when device code cannot be re-run (no GPU at hand): after an edit that is meant to leave existing kernels
alone (a new opt-in path, a host-side change), `--check` proves that every kernel that was verified on the GPU
still compiles to byte-identical instructions.

    python tools/sass_fingerprint.py --save profiles/r1_verified_sass.json      # after a green GPU run
    python tools/sass_fingerprint.py --check profiles/r1_verified_sass.json     # later, on the CPU box
"""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "emdee_b200", "lib", "libemdee.so")


def fingerprints(lib):
    text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, name, buf = {}, None, []

    def flush():
        if name is not None:
            out[name] = {"md5": hashlib.md5("".join(buf).encode()).hexdigest(), "instructions": len(buf)}

    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            # the anonymous-namespace hash changes with every edit of the file: strip it
            name = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "ANON", m.group(1))
            buf = []
        elif name is not None and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
            t = re.sub(r"/\*[0-9a-f]{4,}\*/", "", line)          # instruction address
            t = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", t)             # encoding words
            buf.append(t.strip() + "\n")
    flush()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=LIB)
    ap.add_argument("--save")
    ap.add_argument("--check")
    args = ap.parse_args()
    now = fingerprints(args.lib)
    if args.save:
        json.dump(now, open(args.save, "w"), indent=1, sort_keys=True)
        print(f"{len(now)} kernels -> {args.save}")
    if args.check:
        ref = json.load(open(args.check))
        changed = [k for k in ref if k in now and now[k]["md5"] != ref[k]["md5"]]
        missing = [k for k in ref if k not in now]
        new = [k for k in now if k not in ref]
        print(f"{len(ref)} reference kernels: {len(changed)} changed, {len(missing)} missing; {len(new)} new kernels")
        for k in changed:
            print("  changed:", k[:140])
        for k in missing:
            print("  missing:", k[:140])
        sys.exit(1 if changed or missing else 0)


if __name__ == "__main__":
    main()
