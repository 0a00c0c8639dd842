"""Vendors the UNMODIFIED reference model sources into oracle/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU
box with the snapshot, where /root/reference does not exist) so that bench.py's CPU legs time the reference itself
(cpu_baseline.kind = "reference") instead of the oracle port.

    python oracle/make_ref.py         # in the build container; __graft_entry__.build() runs it when /root/reference exists

Only the files the path needs are copied, byte for byte: models/{__init__,module,drmvsnet}.py and evidential/models.py
(models/drmvsnet.py imports it; the evidential head itself is stubbed at run time, see oracle/ref_runner.py).  Nothing
under oracle/_ref/ is ever committed, and the product never imports it.
"""
import hashlib
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ("models/__init__.py", "models/module.py", "models/drmvsnet.py", "evidential/models.py")


def main():
    if not os.path.isdir(REF):
        print("oracle/make_ref.py: /root/reference is absent here, nothing to do")
        return False
    lines = []
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
        lines.append(f"{hashlib.sha256(open(dst, 'rb').read()).hexdigest()}  {rel}")
    open(os.path.join(DST, "evidential", "__init__.py"), "w").close()
    with open(os.path.join(DST, "SHA256SUMS"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("oracle/_ref: vendored", ", ".join(FILES))
    return True


if __name__ == "__main__":
    main()
