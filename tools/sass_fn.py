"""Extract one kernel's SASS from an object/shared library: python tools/sass_fn.py <file> <substring-of-mangled-name>"""
import subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur, keep = None, []
for line in out.splitlines():
    if "Function :" in line:
        cur = line.split("Function :")[1].strip()
    if cur and sys.argv[2] in cur:
        keep.append(line)
print("\n".join(keep))
