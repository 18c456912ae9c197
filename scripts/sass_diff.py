"""Per-kernel SASS comparison between a git revision and the working tree (GPU-less check that a refactor did
not touch the instruction stream of kernels that were already validated on hardware).

    python scripts/sass_diff.py <git-rev>          # e.g. the last commit that ran on a B200

Compiles every csrc/*.cu of <git-rev> into a temp dir, hashes the instruction text of every kernel (addresses and
encodings stripped) and compares with baton_b200/csrc/build/*.o.  A trailing `, 0` template argument that was
added to an existing kernel (e.g. the CONV mode) is normalised away; new kernels are listed separately."""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math"]


def kernel_hashes(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    out = {}
    for fn in re.split(r"\n\s*Function : ", txt)[1:]:
        name = fn.split("\n", 1)[0].strip()
        ins = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(.*?);", fn)
        out[name] = (hashlib.md5("\n".join(ins).encode()).hexdigest(), len(ins))
    return out


def main():
    rev = sys.argv[1]
    tmp = tempfile.mkdtemp(prefix="sass_diff_")
    subprocess.run("git archive {} baton_b200/csrc | tar -x -C {}".format(rev, tmp), shell=True, check=True, cwd=ROOT)
    old_dir = os.path.join(tmp, "baton_b200", "csrc")
    procs = []
    for f in sorted(os.listdir(old_dir)):
        if f.endswith(".cu"):
            procs.append(subprocess.Popen(["nvcc"] + FLAGS + ["-c", f, "-o", f[:-3] + ".o"], cwd=old_dir,
                                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    for p in procs:
        p.wait()
    total = same = 0
    for f in sorted(os.listdir(old_dir)):
        if not f.endswith(".o"):
            continue
        new_obj = os.path.join(ROOT, "baton_b200", "csrc", "build", f)
        if not os.path.exists(new_obj):
            print("MISSING OBJECT", f)
            continue
        old, new = kernel_hashes(os.path.join(old_dir, f)), kernel_hashes(new_obj)
        norm = {k.replace("ELi0EEEv14", "EEEv14"): v for k, v in new.items()
                if "ELi1EEEv14" not in k and "ELi2EEEv14" not in k}
        for k, v in old.items():
            total += 1
            if norm.get(k) == v:
                same += 1
            else:
                print("DIFF  {:14s} {:5d} -> {:5d}  {}".format(f, v[1], norm.get(k, (0, 0))[1], k[:90]))
        fresh = [k for k in new if k not in old and k.replace("ELi0EEEv14", "EEEv14") not in old]
        if fresh:
            print("NEW   {:14s} {}".format(f, len(fresh)))
    print("kernels in {}: {}   byte-identical instruction stream now: {}".format(rev[:10], total, same))


if __name__ == "__main__":
    main()
