"""Which kernels of an object file changed?  Compares the SASS instruction streams (addresses and .loc lines ignored) of two
builds of one .cu file, kernel by kernel: the no-GPU check that an opt-in variant leaves the default kernels alone.
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -c partition.cu -o before.o   (at the old commit)
    ... edit ...                                                                            -o after.o
    python scripts/sass_diff.py before.o after.o
Kernels are matched by body, not by name: adding a template parameter renames every instantiation."""
import re
import subprocess
import sys


def kernels(obj):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, name, body = {}, None, []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out[name] = body
            name, body = m.group(1), []
        elif name:
            mm = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
            if mm:
                body.append(mm.group(1))
    if name:
        out[name] = body
    return out


def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    bodies = list(b.values())
    changed = [n for n, body in a.items() if body not in bodies]
    print("%d of %d kernels of %s have an identical instruction stream in %s (%d kernels there)" % (len(a) - len(changed), len(a), sys.argv[1], sys.argv[2], len(b)))
    for n in changed:
        print("CHANGED", demangle(n), len(a[n]), "instructions")
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
