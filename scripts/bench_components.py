"""Component timings at the BASELINE.json config sizes (CUDA events, median of 5 after 2 warm-ups).
Writes one JSON line per component; not the headline bench (that is bench.py)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.backends.cuda.matmul.allow_tf32 = False
from mimrl_b200 import _lib as L
from mimrl_b200.model import VMIEstimator, VCMIEstimator, knn_search, prod_knn_sample
from mimrl_b200.mlp_process import MLPEncoder
dev = torch.device("cuda:0")
which = sys.argv[1:] or ["knn", "cubemlp", "vcmi", "concat", "sep_small", "step"]

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))

def out(**kw): print(json.dumps(kw), flush=True)

if "knn" in which:      # config 4: 1M x 128 keys, batch 8192, k = 2..16
    N = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(2)
    Z = torch.randn(N, 128, device=dev, generator=g)
    for k in (2, 16):
        m = 8192 // k
        ids = torch.randperm(N, device=dev, generator=g)[:m]
        ms = timeit(lambda: knn_search(Z, ids, k), reps=3, warm=1)
        out(component="knn_search", N=N, width=128, queries=m, k=k, ms=ms, gflops=2 * 128 * m * N / ms / 1e6,
            key_gbs=4 * 128 * N / ms / 1e6)
    from mimrl_b200.model import KnnPool
    out(component="knn_fit", N=N, width=128, ms=timeit(lambda: KnnPool(Z), 3, 1))
    pool = KnnPool(Z)             # fitted once (per epoch in the training loop), searched many times
    for k in (2, 16):
        m = 8192 // k
        ids = torch.randperm(N, device=dev, generator=g)[:m]
        ms = timeit(lambda: knn_search(pool, ids, k), reps=3, warm=1)
        out(component="knn_search_fitted_pool", N=N, width=128, queries=m, k=k, ms=ms, gflops=2 * 128 * m * N / ms / 1e6,
            key_gbs=4 * 128 * N / ms / 1e6)
    Zl = torch.randn(N, 1, device=dev, generator=g)
    Zi = torch.randint(-3, 4, (N, 1), device=dev, generator=g).float()          # label-like: seven distinct values
    for k in (2, 16):
        ids = torch.randperm(N, device=dev, generator=g)[:8192 // k]
        out(component="knn_search_labels", N=N, width=1, queries=8192 // k, k=k, pool="normal",
            ms=timeit(lambda: knn_search(Zl, ids, k), 3, 1))
        out(component="knn_search_labels", N=N, width=1, queries=8192 // k, k=k, pool="7 distinct values",
            ms=timeit(lambda: knn_search(Zi, ids, k), 3, 1))

if "cubemlp" in which:  # config 5: [1024, 100, 3, 128] -> 50-3-128 -> 10-3-128
    enc = MLPEncoder("gelu", [100, 3, 128], [[50, 3, 128], [10, 3, 128]], [[50, 3, 128], [10, 3, 128]], [0.0] * 3, True,
                     False, [True, True]).to(dev)
    for bs in (128, 1024):
        x = torch.randn(bs, 100, 3, 128, device=dev, requires_grad=True)
        with torch.no_grad():
            fwd = timeit(lambda: enc(x))
        params = list(enc.parameters())
        def fb():
            x.grad = None
            for q in params:           # optimizer.zero_grad(set_to_none=True), as a training step does
                q.grad = None
            enc(x).sum().backward()
        both = timeit(fb)
        alg_bytes = bs * 100 * 384 * 4 * 1.0 + bs * 50 * 384 * 4 * 2 + bs * 10 * 384 * 4          # 330 MB at bs = 1024
        out(component="cubemlp", bs=bs, fwd_ms=fwd, fwd_bwd_ms=both, fwd_gbs=alg_bytes / fwd / 1e6)
        # the same two calls as CUDA graphs (a forward is ~20 launches, fwd+bwd ~70: at these sizes the eager numbers
        # above include host launch time)
        from mimrl_b200.graphs import GraphedCallable
        def fwd_only(xx):
            with torch.no_grad():
                return enc(xx)
        gf = GraphedCallable(fwd_only, [x.detach()])
        fwd_g = timeit(lambda: gf(gf.static_in[0]))          # input already in the graph's static buffer: no copy
        xs = x.detach().clone().requires_grad_(True)
        def fwd_bwd(xx):
            xs.grad = None
            for q in params:
                q.grad = None
            enc(xs).sum().backward()
            return xs.grad
        gb = GraphedCallable(fwd_bwd, [xs.detach()])
        both_g = timeit(lambda: gb(gb.static_in[0]))
        out(component="cubemlp_cuda_graph", bs=bs, fwd_ms=fwd_g, fwd_bwd_ms=both_g, fwd_gbs=alg_bytes / fwd_g / 1e6,
            fwd_frac_of_hbm_peak=alg_bytes / fwd_g / 1e6 / 6546.6)

if "vcmi" in which:
    est = VCMIEstimator(128, 256, 2, "relu", 2, 1.0, "hardtanh").to(dev)
    for bs in (128, 1024, 8192):
        ins = [torch.randn(bs, 128, device=dev, requires_grad=True) for _ in range(6)]
        def fb():
            c, l = est(*ins); (c + l).backward()
        out(component="vcmi_fwd_bwd", bs=bs, ms=timeit(fb))

if "concat" in which:   # config 3 at sizes that fit one GPU quickly
    for bound in ("nwj", "js"):
        est = VMIEstimator("concat", "constant", bound, 128, 256, 128, 2, "relu", 0, 1).to(dev)
        for B in (512, 2048):
            x = torch.randn(B, 128, device=dev, requires_grad=True); y = torch.randn(B, 128, device=dev, requires_grad=True)
            def fb():
                mi, loss = est(x, y); loss.backward()
            ms = timeit(fb, reps=3, warm=1)
            out(component="concat_fwd_bwd", bound=bound, B=B, ms=ms, pairs_per_s=B * B / ms * 1e3)

if "sep_small" in which:  # config 2 sweep
    est = VMIEstimator("separate", "constant", "infonce", 128, 256, 128, 2, "relu", 0, 1).to(dev)
    for B in (128, 512, 2048, 8192, 32768):
        x = torch.randn(B, 128, device=dev, requires_grad=True); y = torch.randn(B, 128, device=dev, requires_grad=True)
        def fb():
            mi, loss = est(x, y); loss.backward()
        ms = timeit(fb)
        out(component="sep_infonce_fwd_bwd", B=B, ms=ms, pairs_per_s=B * B / ms * 1e3)

if "sep_graph" in which or "sep_small" in which:   # config 2, small end of the sweep: one CUDA graph per estimator step
    from mimrl_b200.graphs import GraphedCallable
    for B in (128, 512, 2048, 8192):
        est = VMIEstimator("separate", "constant", "infonce", 128, 256, 128, 2, "relu", 0, 1).to(dev)
        params = list(est.parameters())
        def fb(x, y):
            for p in params:
                p.grad = None
            x.grad = y.grad = None
            mi, loss = est(x, y)
            loss.backward()
            return mi, x.grad, y.grad
        x0 = torch.randn(B, 128, device=dev).requires_grad_(True); y0 = torch.randn(B, 128, device=dev).requires_grad_(True)
        class Wrap:
            def __init__(self):
                self.g = None
            def build(self):
                def fn(xs, ys):
                    xs.requires_grad_(True); ys.requires_grad_(True)
                    return fb(xs, ys)
                self.g = GraphedCallable(fn, [x0, y0])
        w = Wrap(); w.build()
        ms = timeit(lambda: w.g(x0, y0))
        out(component="sep_infonce_fwd_bwd_cuda_graph", B=B, ms=ms, pairs_per_s=B * B / ms * 1e3)

if "step" in which:     # config 1 / 5 shape: one stage-1 + one stage-2 step of the MI/CMI path on synthetic features
    from types import SimpleNamespace
    from mimrl_b200.model import MIHeads
    from mimrl_b200.train_step import FeaturePool, TwoStageStep
    for bs, N in ((128, 1284), (1024, 16326), (8192, 1 << 20)):          # last: BASELINE config 4 pool size
        opt = SimpleNamespace(critic_type="separate", baseline_type="constant", bound_type="infonce", k_neighbor=2,
                              radius=1.0, cmi_last_acticate="hardtanh", d_common=128)
        heads = MIHeads(opt).to(dev)
        enc = torch.nn.Linear(128, 4 * 128).to(dev)          # stand-in for the encoders: features from one projection
        cls = torch.nn.Linear(128, 1).to(dev)
        def features(batch):
            f = enc(batch).view(-1, 4, 128)
            return cls(f[:, 0]), f[:, 0].contiguous(), f[:, 1].contiguous(), f[:, 2].contiguous(), f[:, 3].contiguous()
        main_params = list(enc.parameters()) + list(cls.parameters())
        step = TwoStageStep(heads, features, torch.nn.L1Loss(), torch.optim.Adam(main_params, 1e-4, capturable=True),
                            torch.optim.Adam(heads.parameters(), 1e-4, capturable=True),
                            clip_params=main_params + list(heads.parameters()))
        pool = FeaturePool()
        g = torch.Generator(device="cuda").manual_seed(0)
        pool.C = torch.randn(N, 1, device=dev, generator=g).clamp(-3, 3)
        pool.F, pool.T, pool.A, pool.V = (torch.randn(N, 128, device=dev, generator=g) for _ in range(4))
        batch = torch.randn(bs, 128, device=dev, generator=g)
        labels = torch.randn(bs, device=dev, generator=g).clamp(-3, 3)
        np.random.seed(0)
        def one():
            step.stage1(batch, labels, pool); step.stage2(batch, labels, pool)
        ms = timeit(one, reps=5, warm=2)
        out(component="two_stage_step_mi_cmi", bs=bs, pool=N, ms=ms, steps_per_s=1e3 / ms,
            note="5 VMI + 6 k-NN samplers + 6 VCMI per stage, Adam on both optimisers; encoders replaced by one Linear")
        from mimrl_b200.train_step import GraphedTwoStageStep
        graphed = GraphedTwoStageStep(step, batch, labels, pool, parallel_branches=False)
        def one_g():
            graphed.stage1(batch, labels); graphed.stage2(batch, labels)
        ms = timeit(one_g, reps=5, warm=2)
        out(component="two_stage_step_mi_cmi_cuda_graph", bs=bs, pool=N, ms=ms, steps_per_s=1e3 / ms,
            note="same step, one CUDA graph per stage; k-NN ids drawn on the host before each replay")
        # the eleven estimators of a stage as parallel branches (model.run_branches); eager side streams are not timed:
        # the step is host-bound there (22 ms at bs = 128) and record_stream leaves the allocator fragmented for what follows
        graphed = GraphedTwoStageStep(step, batch, labels, pool)
        ms = timeit(one_g, reps=5, warm=2)
        out(component="two_stage_step_mi_cmi_branches_cuda_graph", bs=bs, pool=N, ms=ms, steps_per_s=1e3 / ms,
            note="one CUDA graph per stage with the estimators as parallel branches")

if "step5" in which or "step" in which:   # config 5 shape: the same two-stage step with the CubeMLP fusion encoder in the loop
    from types import SimpleNamespace
    from mimrl_b200.model import MIHeads
    from mimrl_b200.train_step import FeaturePool, GraphedTwoStageStep, TwoStageStep
    bs, N = 1024, 16326
    opt = SimpleNamespace(critic_type="separate", baseline_type="constant", bound_type="infonce", k_neighbor=2,
                          radius=1.0, cmi_last_acticate="hardtanh", d_common=128)
    heads = MIHeads(opt).to(dev)
    fusion = MLPEncoder("gelu", [100, 3, 128], [[50, 3, 128], [10, 3, 128]], [[50, 3, 128], [10, 3, 128]], [0.0] * 3, True,
                        False, [True, True]).to(dev)                          # README: --d_hiddens 50-3-128=10-3-128
    cls = torch.nn.Linear(128, 1).to(dev)
    def features(batch):                                                      # batch: stacked encoder outputs [bs, 100, 3, 128]
        z = fusion(batch).mean(dim=1)                                         # [bs, 3, 128]
        T_F, A_F, V_F = z[:, 0].contiguous(), z[:, 1].contiguous(), z[:, 2].contiguous()
        F_F = z.mean(dim=1)
        return cls(F_F), F_F, T_F, A_F, V_F
    main_params = list(fusion.parameters()) + list(cls.parameters())
    step = TwoStageStep(heads, features, torch.nn.L1Loss(), torch.optim.Adam(main_params, 1e-4, capturable=True),
                        torch.optim.Adam(heads.parameters(), 1e-4, capturable=True),
                        clip_params=main_params + list(heads.parameters()))
    pool = FeaturePool()
    g = torch.Generator(device="cuda").manual_seed(0)
    pool.C = torch.randn(N, 1, device=dev, generator=g).clamp(-3, 3)
    pool.F, pool.T, pool.A, pool.V = (torch.randn(N, 128, device=dev, generator=g) for _ in range(4))
    batch = torch.randn(bs, 100, 3, 128, device=dev, generator=g)
    labels = torch.randn(bs, device=dev, generator=g).clamp(-3, 3)
    np.random.seed(0)
    def one():
        step.stage1(batch, labels, pool); step.stage2(batch, labels, pool)
    ms = timeit(one, reps=5, warm=2)
    out(component="two_stage_step_cubemlp_mi_cmi", bs=bs, pool=N, ms=ms, steps_per_s=1e3 / ms,
        note="config 5 without BERT/GRU: CubeMLP 50-3-128=10-3-128 on [bs,100,3,128] + 5 VMI + 6 k-NN + 6 VCMI per stage, Adam x2")
    graphed = GraphedTwoStageStep(step, batch, labels, pool, parallel_branches=False)
    def one_g():
        graphed.stage1(batch, labels); graphed.stage2(batch, labels)
    ms = timeit(one_g, reps=5, warm=2)
    out(component="two_stage_step_cubemlp_mi_cmi_cuda_graph", bs=bs, pool=N, ms=ms, steps_per_s=1e3 / ms,
        note="same, one CUDA graph per stage")
    graphed = GraphedTwoStageStep(step, batch, labels, pool)
    ms = timeit(one_g, reps=5, warm=2)
    out(component="two_stage_step_cubemlp_mi_cmi_branches_cuda_graph", bs=bs, pool=N, ms=ms, steps_per_s=1e3 / ms,
        note="same, the estimators of a stage as parallel branches of the graph")
