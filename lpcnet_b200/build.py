"""Builds liblpcnet_b200.so (sm_100a) in-tree with nvcc.  No torch, no JIT cache: the .so travels with the repo
snapshot to the GPU box.  `python -m lpcnet_b200.build` or `lpcnet_b200.build.build()`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "liblpcnet_b200.so")
SOURCES = ["model.cu", "frame_kernels.cu", "sample_kernel.cu", "sample_kernel_f32.cu", "sample_kernel_f32n.cu", "batch_api.cu", "lpcnet_api.cu", "microbench.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-fmad=false",                    # never contract a*b+c: the pinned oracle build uses -ffp-contract=off
         "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off", "-Xptxas", "-v"]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", f) for f in ("lpcnet.h", "lpcnet_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines):
    """Tuning aid: build lpcnet_b200/variants/lib_<name>.so with extra -D flags (e.g. LPCNET_NWC=12)."""
    vdir = os.path.join(HERE, "variants")
    os.makedirs(os.path.join(vdir, "obj_" + name), exist_ok=True)
    objs = []
    for s in SOURCES:
        o = os.path.join(vdir, "obj_" + name, s.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + ["-D%s" % d for d in defines] + ["-c", os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on " + s)
        if s == "sample_kernel.cu":
            print(name, [l.strip() for l in r.stderr.splitlines() if "registers" in l or "spill" in l][-2:])
        objs.append(o)
    so = os.path.join(vdir, "lib_%s.so" % name)
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so] + objs + ["-lcudart"])
    return so


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    objs = []
    os.makedirs(os.path.join(HERE, "_obj"), exist_ok=True)
    log = []
    for s in SOURCES:
        o = os.path.join(HERE, "_obj", s.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on " + s)
        objs.append(o)
    r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + objs + ["-lcudart"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(HERE, "_obj", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
