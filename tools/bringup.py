"""GPU bring-up diagnostics (run under gpurun).  Each stage runs in its own process under a timeout so that a hung
kernel cannot take the whole call down; results are appended to gpurun_out/bringup.log.

    python tools/bringup.py            # all stages
    python tools/bringup.py <stage>    # one stage in-process
"""
import os
import subprocess
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
OUT = os.path.join(REPO, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

STAGES = ["gemm_simt", "attn_simt", "gemm_tc_v0", "gemm_tc_v1", "attn_tc", "model_simt", "model_tc", "timing"]


def _err(a, b):
    a = a.double(); b = b.double()
    return (a - b).abs().max().item(), ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def gemm_cases(path, variant=0):
    import torch
    import torch.nn.functional as F
    from cotr_b200 import capi
    capi.lib().cotr_debug_set_variant(variant)
    g = torch.Generator(device="cpu").manual_seed(0)
    dev = "cuda"
    rnd = lambda *s: torch.randn(*s, generator=g)
    worst = 0.0
    # plain GEMMs (row-major A)
    for (M, N, K) in [(128, 64, 64), (128, 64, 128), (256, 64, 256), (512, 256, 256), (1024, 1024, 256), (512, 256, 1024),
                      (1000, 768, 256), (37, 2, 256), (8192, 64, 256), (300, 3072, 256)]:
        A = rnd(M, K).to(dev); W = rnd(N, K) * 0.1; bias = rnd(N).to(dev)
        ref = (A.double() @ W.to(dev).double().t() + bias.double()).relu()
        out = capi.test_gemm(path, A, W.numpy(), bias=bias, relu=True)
        e = _err(out, ref); worst = max(worst, e[1])
        print(f"  gemm path={path} v={variant} M={M} N={N} K={K}: max {e[0]:.3e} rel {e[1]:.3e}", flush=True)
    # residual + addmat + LayerNorm epilogue
    M, N, K = 700, 256, 1024
    A = rnd(M, K).to(dev); W = rnd(N, K) * 0.05; bias = rnd(N).to(dev); res = rnd(M, N).to(dev)
    gam = (1 + 0.1 * rnd(N)).to(dev); bet = (0.1 * rnd(N)).to(dev)
    ref = F.layer_norm(A.double() @ W.to(dev).double().t() + bias.double() + res.double(), (N,), gam.double(), bet.double(), 1e-5)
    out = capi.test_gemm(path, A, W.numpy(), bias=bias, residual=res, ln=(gam, bet))
    e = _err(out, ref); worst = max(worst, e[1])
    print(f"  gemm+res+LN path={path}: max {e[0]:.3e} rel {e[1]:.3e}", flush=True)
    M, N, K = 1024, 768, 256
    A = rnd(M, K).to(dev); W = rnd(N, K) * 0.05; add = rnd(512, N).to(dev)
    ref = A.double() @ W.to(dev).double().t() + add.double().repeat(2, 1)
    out = capi.test_gemm(path, A, W.numpy(), addmat=add, add_period=512)
    e = _err(out, ref); worst = max(worst, e[1])
    print(f"  gemm+addmat path={path}: max {e[0]:.3e} rel {e[1]:.3e}", flush=True)
    # implicit-GEMM convolutions (NHWC)
    for (n, H, C, Co, k, s, pd) in [(2, 16, 64, 64, 3, 1, 1), (2, 32, 128, 128, 3, 2, 1), (2, 16, 256, 256, 3, 2, 1), (2, 32, 256, 512, 1, 2, 0)]:
        x = rnd(n, C, H, H); w = rnd(Co, C, k, k) * 0.05; bias = rnd(Co).to(dev)
        ref = F.conv2d(x.to(dev).double(), w.to(dev).double(), bias.double(), stride=s, padding=pd).permute(0, 2, 3, 1).reshape(-1, Co)
        OH = (H + 2 * pd - k) // s + 1
        xn = x.permute(0, 2, 3, 1).contiguous().to(dev)
        wk = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
        out = capi.test_gemm(path, xn, wk.numpy(), bias=bias, a_mode=1, M=n * OH * OH,
                             conv=dict(H=H, W=H, C=C, OH=OH, OW=OH, KH=k, KW=k, stride=s, pad=pd))
        e = _err(out, ref); worst = max(worst, e[1])
        print(f"  conv path={path} C={C}->{Co} k={k} s={s}: max {e[0]:.3e} rel {e[1]:.3e}", flush=True)
    # stem: 7x7/2 over the NCHW canvas, halves as separate images
    img = rnd(1, 3, 256, 512); w = rnd(64, 3, 7, 7) * 0.1; bias = rnd(64).to(dev)
    halves = torch.cat([img[..., :256], img[..., 256:]], 0)
    ref = F.conv2d(halves.to(dev).double(), w.to(dev).double(), bias.double(), stride=2, padding=3).relu().permute(0, 2, 3, 1).reshape(-1, 64)
    wk = w.permute(0, 2, 3, 1).reshape(64, -1).contiguous()
    out = capi.test_gemm(path, img.to(dev), wk.numpy(), bias=bias, relu=True, a_mode=2, M=2 * 128 * 128,
                         conv=dict(H=256, W=256, C=3, OH=128, OW=128, KH=7, KW=7, stride=2, pad=3))
    e = _err(out, ref); worst = max(worst, e[1])
    print(f"  stem path={path}: max {e[0]:.3e} rel {e[1]:.3e}", flush=True)
    print(f"WORST rel {worst:.3e}", flush=True)


def tc_precision():
    """Where does the tcgen05 GEMM error come from?  fp16-exact operands isolate the accumulate; variants: 0 = split
    accumulators (default), 2 = hi*hi only, 4 = everything chained on one accumulator."""
    import torch
    from cotr_b200 import capi
    g = torch.Generator(device="cpu").manual_seed(0)
    for K in (64, 256, 1024, 2304):
        M, N = 256, 128
        A = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g)
        A16 = A.half().float(); W16 = (W.half().float())          # exactly representable: lo terms vanish
        for name, a, w, variants in (("fp16-exact", A16, W16, (2, 0, 4)), ("fp32", A, W, (0, 4)), ("fp32 positive", A.abs(), W.abs(), (0, 4))):
            ref = a.cuda().double() @ w.cuda().double().t()
            for v in variants:
                capi.lib().cotr_debug_set_variant(v)
                out = capi.test_gemm(0, a.cuda(), w.numpy())
                e = _err(out, ref)
                bias = ((out.double() - ref) / ref.abs().clamp_min(1e-9)).mean().item()
                print(f"  K={K:5d} {name:14s} variant={v}: rel {e[1]:.3e}  mean signed rel err {bias:+.3e}", flush=True)
            out = capi.test_gemm(1, a.cuda(), w.numpy())
            print(f"  K={K:5d} {name:14s} simt fp32 : rel {_err(out, ref)[1]:.3e}", flush=True)
    capi.lib().cotr_debug_set_variant(0)


def gemm_timeline():
    """clock64() timeline of one CTA of the tcgen05 GEMM (cotr_debug_set_timestamps).  Slots: 1 setup done; 2 first A
    fetch issued; 3+2i loader got stage i; 4+2i loader published stage i; 20 accumulator ready (epilogue start);
    21 epilogue done; 24+2i MMA saw A(i); 25+2i MMA issued + committed stage i; 41 final commit; 44+i TMA issued
    stage i; 60 teardown done."""
    import ctypes
    import torch
    from cotr_b200 import capi
    g = torch.Generator(device="cpu").manual_seed(0)
    for (M, N, K, ln) in [(1024, 256, 256, False), (512, 1024, 256, False), (512, 1024, 256, "a_ln"), (512, 256, 1024, False),
                          (512, 256, 1024, "res_ln"), (512, 256, 2304, False), (512, 256, 1024, True)]:
        A = torch.randn(M, K, generator=g).cuda(); W = torch.randn(N, K, generator=g) * 0.1
        bias = torch.randn(N, generator=g).cuda()
        ln_args = (torch.ones(256).cuda(), torch.zeros(256).cuda()) if ln else None
        res = torch.randn(M, N, generator=g).cuda() if ln == "res_ln" else None
        ts = torch.zeros(64 * 1024, dtype=torch.int64, device="cuda")
        for rep in range(3):
            ts.zero_()
            capi.lib().cotr_debug_set_timestamps(ctypes.c_void_p(ts.data_ptr()))
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            capi.test_gemm(0, A, W.numpy(), bias=bias, ln=ln_args, a_ln=(ln == "a_ln"), res_ln=(ln == "res_ln"), residual=res)
        capi.lib().cotr_debug_set_timestamps(None)
        t = ts.cpu().view(-1, 64)
        for cta in (0, 1):
            r = t[cta].tolist()
            print(f"  M={M} N={N} K={K} ln={ln} cta{cta}: setup {r[1]} fetch0 {r[2]} | loader got/pub " +
                  " ".join(f"{r[3 + 2 * i]}/{r[4 + 2 * i]}" for i in range(min(8, (K + 63) // 64))) +
                  f" | mma sawA/issued " + " ".join(f"{r[24 + 2 * i]}/{r[25 + 2 * i]}" for i in range(min(8, (K + 63) // 64))) +
                  f" | tma " + " ".join(str(r[44 + i]) for i in range(min(8, (K + 63) // 64))) +
                  f" | final commit {r[41]} acc ready {r[20]} | epi chunk0 acc/apply/emit {r[30]}/{r[31]}/{r[32]} chunk1 {r[34]}/{r[35]}/{r[36]} before drain {r[38]}"
                  f" epi done {r[21]} end {r[60]} | stats first chunk {r[50]} done {r[51]}", flush=True)


def backbone_timeline():
    """The same timeline for the backbone's heavy shapes: stem (7x7 on the fp32 canvas), layer1 conv3 (K=64 with a
    residual and 8 MB of output), layer1 3x3, layer2 conv1."""
    import ctypes
    import torch
    from cotr_b200 import capi
    g = torch.Generator(device="cpu").manual_seed(0)
    rnd = lambda *s: torch.randn(*s, generator=g)
    ts = torch.zeros(64 * 1024, dtype=torch.int64, device="cuda")

    def show(tag, nchunks, ctas=(0, 1, 200)):
        t = ts.cpu().view(-1, 64)
        for cta in ctas:
            r = t[cta].tolist()
            if r[60] == 0:
                continue
            n = min(6, nchunks)
            print(f"  {tag} cta{cta}: setup {r[1]} fetch0 {r[2]} | loader got/pub " +
                  " ".join(f"{r[3 + 2 * i]}/{r[4 + 2 * i]}" for i in range(n)) +
                  " | mma sawA/issued " + " ".join(f"{r[24 + 2 * i]}/{r[25 + 2 * i]}" for i in range(n)) +
                  f" | final commit {r[41]} acc ready {r[20]} | chunk0 acc/apply/emit {r[30]}/{r[31]}/{r[32]} chunk1 {r[34]}/{r[35]}/{r[36]}"
                  f" before drain {r[38]} epi done {r[21]} end {r[60]} | cluster barrier in/out {r[22]}/{r[23]}", flush=True)

    def run(tag, nchunks, fn):
        for rep in range(3):
            ts.zero_()
            capi.lib().cotr_debug_set_timestamps(ctypes.c_void_p(ts.data_ptr()))
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            fn()
        capi.lib().cotr_debug_set_timestamps(None)
        show(tag, nchunks)

    img = rnd(1, 3, 256, 512).cuda(); w = (rnd(64, 3, 7, 7) * 0.1).permute(0, 2, 3, 1).reshape(64, -1).contiguous().numpy()
    bias64 = rnd(64).cuda()
    run("stem 32768x64x147", 3, lambda: capi.test_gemm(0, img, w, bias=bias64, relu=True, a_mode=2, M=2 * 128 * 128,
                                                        conv=dict(H=256, W=256, C=3, OH=128, OW=128, KH=7, KW=7, stride=2, pad=3)))
    A = rnd(8192, 64).cuda(); W = (rnd(256, 64) * 0.1).numpy(); b = rnd(256).cuda(); res = rnd(8192, 256).cuda()
    run("l1.conv3 8192x256x64 +res+relu", 1, lambda: capi.test_gemm(0, A, W, bias=b, residual=res, relu=True))
    A2 = rnd(8192, 256).cuda(); W2 = (rnd(64, 256) * 0.1).numpy()
    run("l1.conv1 8192x64x256 relu", 4, lambda: capi.test_gemm(0, A2, W2, bias=bias64, relu=True))
    x = rnd(2, 64, 64, 64).permute(0, 2, 3, 1).contiguous().cuda(); w3 = (rnd(64, 64, 3, 3) * 0.05).permute(0, 2, 3, 1).reshape(64, -1).contiguous().numpy()
    run("l1.conv2 3x3 8192x64x576", 9, lambda: capi.test_gemm(0, x, w3, bias=bias64, relu=True, a_mode=1, M=8192,
                                                               conv=dict(H=64, W=64, C=64, OH=64, OW=64, KH=3, KW=3, stride=1, pad=1)))
    x3 = rnd(2, 16, 16, 256).cuda(); w33 = (rnd(256, 256, 3, 3) * 0.05).permute(0, 2, 3, 1).reshape(256, -1).contiguous().numpy()
    b256 = rnd(256).cuda()
    # split-K 4: grid (4, 8, 4); CTA 0 is a leader, CTA 32 the first peer (z = 1)
    run("l3.conv2 3x3 512x256x2304 (split-K)", 9, lambda: capi.test_gemm(0, x3, w33, bias=b256, relu=True, a_mode=1, M=512,
                                                                            conv=dict(H=16, W=16, C=256, OH=16, OW=16, KH=3, KW=3, stride=1, pad=1)))
    show("   ... peers", 9, ctas=(32, 64))
    A4 = rnd(512, 1024).cuda(); W4 = (rnd(256, 1024) * 0.05).numpy()
    run("l3.conv1 512x256x1024 (split-K)", 4, lambda: capi.test_gemm(0, A4, W4, bias=b256, relu=True))
    show("   ... peers", 4, ctas=(32, 64))
    A3 = rnd(2048, 128).cuda(); W3 = (rnd(512, 128) * 0.1).numpy(); b3 = rnd(512).cuda(); res3 = rnd(2048, 512).cuda()
    run("l2.conv3 2048x512x128 +res+relu", 2, lambda: capi.test_gemm(0, A3, W3, bias=b3, residual=res3, relu=True))


def attn_timeline():
    """clock64() timeline of the attention kernel.  Slots: 1 setup; 2 after griddepcontrol.wait; 3 staging issued;
    4 S ready (softmax start); 5 row max done; 6+c P chunk c handed over; 14 O ready; 15 O stored; MMA thread: 20 Q/K
    landed; 21 S issued; 22 V landed; 24+2c / 25+2c P chunk c seen / its MMAs issued; 60 end."""
    import ctypes
    import torch
    from cotr_b200 import capi
    g = torch.Generator(device="cpu").manual_seed(1)
    ts = torch.zeros(64 * 1024, dtype=torch.int64, device="cuda")
    for (nq, npairs) in [(512, 1), (1024, 1)]:
        q = torch.randn(npairs * nq, 256, generator=g).cuda()
        k = torch.randn(npairs * 512, 256, generator=g).cuda()
        v = torch.randn(npairs * 512, 256, generator=g).cuda()
        for rep in range(3):
            ts.zero_()
            capi.lib().cotr_debug_set_timestamps(ctypes.c_void_p(ts.data_ptr()))
            capi.test_attention(0, q, k, v, nq, npairs)
        capi.lib().cotr_debug_set_timestamps(None)
        t = ts.cpu().view(-1, 64)
        for cta in (0, 5):
            r = t[cta].tolist()
            print(f"  attention nq={nq} cta{cta}: setup {r[1]} pdl {r[2]} staged {r[3]} | mma: qk landed {r[20]} S issued {r[21]} v landed {r[22]} | "
                  f"softmax: S ready {r[4]} max done {r[5]} P chunks " + " ".join(str(r[6 + c]) for c in range(8)) +
                  " | mma saw/issued " + " ".join(f"{r[24 + 2 * c]}/{r[25 + 2 * c]}" for c in range(8)) +
                  f" | O ready {r[14]} stored {r[15]} end {r[60]}", flush=True)


def forward_trace():
    """%globaltimer trace of one graph-replayed forward (B=1, Q=1024): per tcgen05 launch, CTA 0's kernel entry, the
    return of griddepcontrol.wait and the CTA exit, in microseconds since the first launch (variant bit 17)."""
    import ctypes
    import torch
    from cotr_b200 import capi
    from cotr_b200.models import build_model
    from cotr_b200.utils import synthetic as fixtures
    sd = fixtures.make_state_dict(0)
    model = build_model(None)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.cuda().eval()
    img, q = fixtures.make_inputs(1, 1, 1024)
    img = torch.from_numpy(img).cuda(); q = torch.from_numpy(q).cuda()
    n_launch = int(os.environ.get("COTR_TRACE_N", "112"))
    ts = torch.zeros(n_launch * 256 * 64, dtype=torch.int64, device="cuda")
    capi.lib().cotr_debug_set_variant((1 << 17) | int(os.environ.get("COTR_TRACE_VARIANT", "0")))
    for rep in range(4):                      # eager, capture, replay, replay
        capi.lib().cotr_debug_set_timestamps(ctypes.c_void_p(ts.data_ptr()))
        model(img, q)
        torch.cuda.synchronize()
    capi.lib().cotr_debug_set_timestamps(None)
    capi.lib().cotr_debug_set_variant(0)
    t = ts.cpu().view(n_launch, 256, 64)
    t0 = None
    prev_end = None
    rows = []
    for i in range(n_launch):
        ctas = t[i]
        # %globaltimer is read once per CTA, at exit (slot 62), next to the exit cycle stamp (slot 60): a second read
        # shortly after a first one stalls the reading thread for microseconds (measured), so entry and wait-done times
        # are reconstructed from the SM cycle counter
        ctas = ctas.clone()
        ctas[:, 61] = torch.where(ctas[:, 62] > 0, ctas[:, 62] - (ctas[:, 60].double() / 1.965).long(), ctas[:, 61])
        used = ctas[:, 62] > 0
        if not used.any():
            continue
        # the wait-done time comes from CTA 0's cycle counter: a second %globaltimer read shortly after the first
        # stalls the reading thread for several microseconds (measured), so kernels read it at entry and exit only
        start = ctas[used, 61].min().item(); first = ctas[0, 61].item(); waited = first + int(ctas[0, 2].item() / 1.965)
        end = ctas[used, 62].max().item(); end0 = ctas[0, 62].item()
        if t0 is None:
            t0 = start
        rows.append((i, int(used.sum()), (start - t0) / 1e3, (waited - t0) / 1e3, (end0 - t0) / 1e3, (end - t0) / 1e3))
    for i in [int(x) for x in os.environ.get("COTR_TRACE_DETAIL", "").split(",") if x]:
        ctas = t[i].clone()
        ctas[:, 61] = torch.where(ctas[:, 62] > 0, ctas[:, 62] - (ctas[:, 60].double() / 1.965).long(), ctas[:, 61])
        for c in range(256):
            if ctas[c, 61] > 0 and (c < 6 or c % 16 == 0):
                print(f"    launch {i} cta {c:3d}: start {(ctas[c, 61].item() - t0) / 1e3:8.2f} wait done {(ctas[c, 61].item() + ctas[c, 2].item() / 1.965 - t0) / 1e3:8.2f} "
                      f"end {(ctas[c, 62].item() - t0) / 1e3:8.2f}   cycles: setup {ctas[c, 1].item()} wait {ctas[c, 2].item()} end {ctas[c, 60].item()}", flush=True)
    for i in [int(x) for x in os.environ.get("COTR_TRACE_SLOTS", "").split(",") if x]:
        for c in (0, 1, 9):
            r = t[i][c].tolist()
            if r[62] > 0:
                print(f"    launch {i} cta {c} cycle stamps: " + " ".join(f"{k}:{r[k]}" for k in range(1, 61) if r[k] > 0), flush=True)
    for (i, n, start, waited, end0, end) in rows:
        gap = "" if prev_end is None else f" gap after prev end {start - prev_end:+6.2f}"
        print(f"  launch {i:3d} ctas {n:3d}: first CTA start {start:8.2f}  cta0 wait done {waited:8.2f}  cta0 end {end0:8.2f}  last CTA end {end:8.2f} us{gap}", flush=True)
        prev_end = end


def launch_profile():
    """Per-launch CUDA-event durations of one eager forward (B=1, Q=1024), library profiler."""
    import torch
    from cotr_b200.models import build_model
    from cotr_b200.utils import synthetic as fixtures
    sd = fixtures.make_state_dict(0)
    model = build_model(None)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.cuda().eval()
    img, q = fixtures.make_inputs(1, int(os.environ.get("COTR_PROFILE_B", "1")), int(os.environ.get("COTR_PROFILE_Q", "1024")))
    img = torch.from_numpy(img).cuda(); q = torch.from_numpy(q).cuda()
    nat = model.native()
    for _ in range(3):
        model(img, q)
    import collections
    acc = collections.OrderedDict()
    reps = 5
    for _ in range(reps):
        nat.profile_begin(1024)
        model(img, q)
        for i, (name, M, N, K, ms) in enumerate(nat.profile_end()):
            key = (i, name, M, N, K)
            acc[key] = acc.get(key, 0.0) + ms
    total = 0.0
    by_kind = collections.defaultdict(float)
    for (i, name, M, N, K), ms in acc.items():
        us = ms / reps * 1e3
        total += us
        by_kind[name] += us
        print(f"  {i:3d} {name:14s} M={M:6d} N={N:5d} K={K:5d}  {us:7.1f} us", flush=True)
    print(f"  total {total:.0f} us; by kind: " + ", ".join(f"{k} {v:.0f}" for k, v in by_kind.items()), flush=True)


def attn_cases(path):
    import torch
    from cotr_b200 import capi
    g = torch.Generator(device="cpu").manual_seed(1)
    dev = "cuda"
    for (nq, npairs, gain) in [(512, 1, 1.0), (1024, 2, 2.0), (100, 3, 1.0), (257, 1, 3.0), (1, 4, 1.0)]:
        q = (torch.randn(npairs * nq, 256, generator=g) * gain).to(dev)
        k = torch.randn(npairs * 512, 256, generator=g).to(dev)
        v = torch.randn(npairs * 512, 256, generator=g).to(dev)
        qh = q.double().view(npairs, nq, 8, 32).transpose(1, 2)
        kh = k.double().view(npairs, 512, 8, 32).transpose(1, 2)
        vh = v.double().view(npairs, 512, 8, 32).transpose(1, 2)
        ref = (torch.softmax(qh @ kh.transpose(-1, -2), -1) @ vh).transpose(1, 2).reshape(npairs * nq, 256)
        out = capi.test_attention(path, q, k, v, nq, npairs)
        e = _err(out, ref)
        print(f"  attention path={path} nq={nq} pairs={npairs} gain={gain}: max {e[0]:.3e} rel {e[1]:.3e}", flush=True)


def model_case(path):
    import torch
    from cotr_b200.models import build_model
    from oracle import cotr_oracle          # bring-up check against the CPU oracle
    from cotr_b200.utils import synthetic as fixtures
    sd = fixtures.make_state_dict(0)
    img, q = fixtures.make_inputs(1, 1, 1024)
    model = build_model(None)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.cuda().eval()
    model.native().set_gemm_path(path)
    pred = model(torch.from_numpy(img).cuda(), torch.from_numpy(q).cuda())["pred_corrs"].cpu()
    o32, inter = cotr_oracle.forward(sd, img, q, torch.float32, return_intermediates=True)
    o64 = cotr_oracle.forward(sd, img, q, torch.float64)
    nat = model.native()
    feat = torch.from_numpy(nat.debug_read("feat", 2 * 16 * 16 * 1024)).view(2, 16, 16, 1024)
    feat = torch.cat([feat[0], feat[1]], dim=1).permute(2, 0, 1)[None]          # (1,1024,16,32)
    src = torch.from_numpy(nat.debug_read("src", 512 * 256)).view(1, 512, 256)
    mem = torch.from_numpy(nat.debug_read("mem", 512 * 256)).view(1, 512, 256)
    hs = torch.from_numpy(nat.debug_read("hs", 1024 * 256)).view(1, 1024, 256)
    pos = torch.from_numpy(nat.debug_read("pos", 512 * 256)).view(512, 256)
    for name, mine, ref in (("pos", pos, inter["pos"]), ("feat", feat, inter["feat"]), ("src", src, inter["src"]),
                            ("mem", mem, inter["mem"]), ("hs", hs, inter["hs"])):
        e = _err(mine, ref)
        print(f"  {name}: max {e[0]:.3e} rel {e[1]:.3e}", flush=True)
    print(f"  pred vs oracle fp32: max {(pred - o32).abs().max().item():.3e}; vs fp64: {(pred.double() - o64).abs().max().item():.3e}; "
          f"oracle32 vs 64: {(o32.double() - o64).abs().max().item():.3e}; launches {nat.last_launch_count()}", flush=True)
    gold = np.load(os.path.join(REPO, "tests", "golden", "model_b1_q1024.npz"))
    print(f"  pred vs golden ref fp32: {np.abs(pred.numpy() - gold['ref_pred_fp32']).max():.3e}", flush=True)


def timing():
    import torch
    from cotr_b200.models import build_model
    from cotr_b200.utils import synthetic as fixtures
    sd = fixtures.make_state_dict(0)
    model = build_model(None)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.cuda().eval()
    for path in (1, 0):
        model.native().set_gemm_path(path)
        for (B, Q) in ((1, 1024), (8, 1024), (32, 1), (1, 16384)):
            img, q = fixtures.make_inputs(1, B, Q)
            img = torch.from_numpy(img).cuda(); q = torch.from_numpy(q).cuda()
            for _ in range(3):
                model(img, q)
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            n = 10
            t0.record()
            for _ in range(n):
                model(img, q)
            t1.record(); torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / n
            print(f"  path={path} B={B} Q={Q}: {ms:.3f} ms/forward  {B * Q / ms * 1e3:.0f} q/s  launches {model.native().last_launch_count()}", flush=True)


def run_stage(name):
    if os.environ.get("COTR_B200_LIB"):          # dev: run a stage against another build (tools/ab_libs.py)
        from cotr_b200 import capi
        capi.LIB_PATH = os.environ["COTR_B200_LIB"]
    if name == "gemm_simt":
        gemm_cases(1)
    elif name == "attn_simt":
        attn_cases(1)
    elif name == "gemm_tc_v0":
        gemm_cases(0, 0)
    elif name == "gemm_tc_v1":
        gemm_cases(0, 1)
    elif name == "attn_tc":
        attn_cases(0)
    elif name == "model_simt":
        model_case(1)
    elif name == "model_tc":
        model_case(0)
    elif name == "timing":
        timing()
    elif name == "launch_profile":
        launch_profile()
    elif name == "gemm_timeline":
        gemm_timeline()
    elif name == "forward_trace":
        forward_trace()
    elif name == "attn_timeline":
        attn_timeline()
    elif name == "backbone_timeline":
        backbone_timeline()
    elif name == "tc_precision":
        tc_precision()
    else:
        raise SystemExit(f"unknown stage {name}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_stage(sys.argv[1])
        sys.exit(0)
    log = open(os.path.join(OUT, "bringup.log"), "a")
    for st in STAGES:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), st], capture_output=True, text=True, timeout=240)
            tail = r.stdout + ("\nSTDERR:\n" + r.stderr[-3000:] if r.returncode else "")
            status = f"rc={r.returncode}"
        except subprocess.TimeoutExpired as e:
            tail = (e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")) + "\nTIMEOUT"
            status = "TIMEOUT"
        msg = f"=== {st}: {status} ({time.time() - t0:.1f}s)\n{tail}\n"
        print(msg, flush=True)
        log.write(msg); log.flush()
