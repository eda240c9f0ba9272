"""A/B timing of two builds of libcotr_b200.so on the same GPU box (box-to-box variance is several percent, larger
than most kernel changes, so candidates are compared against a baseline build inside one gpurun call).

    python tools/ab_libs.py build <git-rev> <name>     # here (no GPU): csrc of <git-rev> -> gpurun_out/../ab/<name>.so
    python tools/ab_libs.py run <name-or-path> ...     # on the GPU: alternate the libraries, fresh process each
    python tools/ab_libs.py one <path>                 # (internal) one measurement

"current" names the in-tree library.  Libraries built by `build` live under cotr_b200/lib/ab/ (git-ignored with the
rest of lib/, but they travel with the gpurun snapshot).
"""
import os
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
AB_DIR = os.path.join(REPO, "cotr_b200", "lib", "ab")


def lib_path(name):
    if name == "current":
        return os.path.join(REPO, "cotr_b200", "lib", "libcotr_b200.so")
    return name if os.path.sep in name else os.path.join(AB_DIR, name + ".so")


def build(rev, name):
    from cotr_b200 import build as b
    os.makedirs(AB_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(f"git -C {REPO} archive {rev} cotr_b200/csrc include | tar -x -C {tmp}", shell=True, check=True)
        csrc = os.path.join(tmp, "cotr_b200", "csrc")
        objs = []
        procs = []
        for src in [f for f in os.listdir(csrc) if f.endswith(".cu")]:
            obj = os.path.join(tmp, src.replace(".cu", ".o"))
            flags = [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
            procs.append(subprocess.Popen([b._nvcc(), *flags, "-c", os.path.join(csrc, src), "-o", obj]))
            objs.append(obj)
        for p in procs:
            if p.wait() != 0:
                raise SystemExit("nvcc failed")
        out = lib_path(name)
        subprocess.run([b._nvcc(), "-shared", "-cudart", "static", "-o", out, *objs], check=True)
        print(out)


def one(path):
    os.environ["COTR_B200_ALLOW_OLD_LIB"] = "1"
    import torch
    from cotr_b200 import capi
    capi.LIB_PATH = path
    from cotr_b200.models import build_model
    from cotr_b200.utils import synthetic as fixtures
    sd = fixtures.make_state_dict(0)
    model = build_model(None)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.cuda().eval()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    out = []
    shapes = ((1, 1024, 60),) if os.environ.get("COTR_AB_QUICK") else ((1, 1024, 60), (8, 1024, 12), (1, 16384, 12))
    for (B, Q, n) in shapes:
        img, q = fixtures.make_inputs(1, B, Q)
        img = torch.from_numpy(img).cuda(); q = torch.from_numpy(q).cuda()
        for _ in range(5):
            model(img, q)
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            flush.zero_()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); model(img, q); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        out.append(f"B={B} Q={Q}: median {ts[len(ts) // 2]:.4f} ms (min {ts[0]:.4f})")
    print(" | ".join(out), flush=True)


def run(names, rounds=int(os.environ.get("COTR_AB_ROUNDS", "3"))):
    for r in range(rounds):
        for name in names:
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "one", lib_path(name)], capture_output=True, text=True)
            line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else res.stderr.strip()[-300:]
            print(f"  round {r} {name:12s}: {line}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "one":
        one(sys.argv[2])
    else:
        run(sys.argv[2:])
