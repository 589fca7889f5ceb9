#!/usr/bin/env python
"""Are two builds of an object file the same GPU code?  Compares the SASS instruction streams (addresses and encodings
stripped) of every kernel that exists in both, by demangled name with defaulted trailing `false` template arguments
dropped.  Used to show that adding the fp16-operand template variant left every existing tcgen05 kernel untouched when no
GPU was available to re-run the parity tests:
    python profiles/sass_identity.py old/conv_tc.o new/conv_tc.o"""
import re
import subprocess
import sys


def funcs(obj):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    out, cur, buf = {}, None, []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                out[cur] = buf
            cur, buf = m.group(1), []
        elif cur is not None:
            t = re.sub(r"/\*[0-9a-fx ]+\*/", "", line).strip()
            if t and not t.startswith("."):
                buf.append(t)
    if cur:
        out[cur] = buf
    names = list(out)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    norm = lambda n: re.sub(r"\(.*", "", re.sub(r"(, false)+>", ">", n))
    return {norm(d): out[n] for n, d in zip(names, dem)}


def main():
    a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
    same = [k for k in a if a[k] == b.get(k)]
    print("%d kernels in the old object, %d identical in the new one, %d changed or missing, %d new" %
          (len(a), len(same), len(a) - len(same), len([k for k in b if k not in a])))
    for k in a:
        if k not in same:
            print("  CHANGED:", k)
    for k in b:
        if k not in a:
            print("  new:", k)
    return 0 if len(same) == len(a) else 1


if __name__ == "__main__":
    sys.exit(main())
