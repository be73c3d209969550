"""bench.py -- headline benchmark of the B200-native instance-embedding hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|infer]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workloads (BASELINE.json):
  train (default, configs[1] / configs[2]):  one training step of the designed ReSeg path -- backbone, 2 x ReNet
        (GRU scan kernels), multi-head attention (tcgen05 kernel), heads, DiscriminativeLoss + CE + Dice,
        backward, clip, Adadelta -- on CVPPP-shaped synthetic batches of 16 images PER GPU (weak scaling;
        8 GPUs = the global batch 128 of configs[2]); data parallel = one flat NCCL gradient all-reduce.
  infer (configs[0]): pred.py: one 530x500 image -> resize 256x256 -> net -> softmax -> device clustering
        (k=16, n_init=35, max_iter=500) -> masks up-sampled to 530x500; N GPUs = N independent replicas.
Prints ONE JSON line (rank 0).  `value` = whole-job images/s with inputs resident in HBM, `e2e` = the same
through the public API with pinned HOST buffers (H2D + D2H inside the timed region).
`--impl reference` times the reference's CPU implementation of the same workload (oracle/model_ref.py:
nn.GRU ReNet, the reference's MultiHeadAttention math, its broadcast discriminative-loss graph, real
scikit-learn KMeans, cv2) on the box's host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PER_GPU_BATCH = 16
NET_H = NET_W = 256
RAW_H, RAW_W = 530, 500
C_EMB, K_MAX, N_OBJ = 24, 32, 16


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops_sustained"]), src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- synthetic data
def train_batch(seed, bs):
    from isa_b200 import synth
    d = synth.batch(seed, bs, 3, NET_H, NET_W, K_MAX)          # d["emb"] doubles as the (b,3,H,W) image tensor
    labels = d["labels"]
    sem = np.stack([(labels == 255), (labels != 255)], 1).astype(np.int64)   # one-hot (b,2,H,W) int64 (dataset.py:354-376)
    ins = synth.onehot(labels, K_MAX, np.int64)                               # one-hot (b,32,H,W) int64
    return d["emb"], sem, ins, labels, d["n_objects"].astype(np.int32)


# ---------------------------------------------------------------------------------------------- our arm
def algorithmic(name, cfg):
    """(bound, units per launch) used for roofline.achieved; formulas documented in DESIGN.md section 4."""
    bs, HW = cfg["bs"], NET_H * NET_W
    tok = cfg["bs"] * (NET_H // 4) * (NET_W // 4)
    n = 100
    if name == "isa_disc_loss_fwd":
        return "hbm", bs * HW * (4 * C_EMB + cfg["tgt_bytes"]) + bs * K_MAX * C_EMB * 4
    if name == "isa_disc_loss_bwd":     # emb read + grad write + the 1 B/pixel label map distilled by the forward call
        return "hbm", bs * HW * (8 * C_EMB + 1)
    if name == "isa_gru_scan_fwd":      # gx read + h write + stash write, both directions
        return "hbm", tok * 2 * (3 * n + n + 4 * n) * 4
    if name == "isa_gru_scan_bwd":      # dout + out + stash read, dgx + dghn write
        return "hbm", tok * 2 * (n + n + 4 * n + 3 * n + n) * 4
    if name == "isa_attention_fwd":     # 4 * BH * Lq * Lk * d flops
        L = (NET_H // 4) * (NET_W // 4)
        return "tensor", 4.0 * (2 * bs) * L * L * 12
    if name == "isa_attention_bwd":
        L = (NET_H // 4) * (NET_W // 4)
        return "tensor", 10.0 * (2 * bs) * L * L * 12
    # channels-last epilogues: 9 conv / transposed-conv outputs per step, E = bs*HW*315 elements in total
    # (stage1 2x64, stage2 2x128/4, stage3 3x256/16, up1 100/4, up2 50 channels per full-resolution pixel)
    if name == "isa_bias_act_fwd":      # read + write in place, averaged over the 9 calls
        return "hbm", bs * HW * 315 * 8 / 9.0
    if name == "isa_bias_act_bwd":      # gy read + y read + gx write
        return "hbm", bs * HW * 315 * 12 / 9.0
    if name in ("isa_pixel_heads_fwd", "isa_pixel_heads_bwd", "isa_pixel_heads_wgrad"):
        # (50 + 64) source channels and (2 + 24) output planes per pixel, each touched once
        return "hbm", bs * HW * (114 + 26) * 4
    if name == "isa_add_layernorm_fwd":
        return "hbm", tok * 24 * 4 * 3
    if name == "isa_add_layernorm_bwd":
        return "hbm", tok * 24 * 4 * 4
    if name == "isa_split_bf16x3":      # fp32 read + 3 bf16 parts written, averaged over the calls of one step
        return "hbm", cfg.get("split_bytes_per_call", 0.0)
    return None, 0.0


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel behind `name`, from the committed
    `ncu --set full` captures (profiles/r1_ncu_traffic.json); None when no capture exists for it."""
    p = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(name, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def roofline_entry(name, n_calls, total_ms, cfg, peaks):
    bound, units = algorithmic(name, cfg)
    if bound is None or n_calls == 0 or total_ms <= 0:
        return None
    per_launch_s = total_ms / n_calls * 1e-3
    if bound == "hbm":
        ach, peak, unit = units / per_launch_s / 1e9, peaks["hbm"], "GB/s"
    else:
        ach, peak, unit = units / per_launch_s / 1e12, peaks["bf16"], "TFLOP/s"
    return {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
            "traffic": ncu_traffic(name), "peak_source": peaks["src"], "launches_timed": n_calls, "avg_launch_us": per_launch_s * 1e6}


METRIC = "images/sec (pred.py inference, train step at 1/2/4/8 B200) vs host-CPU ref"
INFER_WORKLOAD = ("pred.py inference: 530x500 RGB -> 256x256 -> ReSeg path -> softmax -> device k-means "
                  "(k=16, n_init=35, max_iter=500, seed 0) -> masks at 530x500; batch 1 per GPU, replicas only")
TRAIN_WORKLOAD = ("CVPPP-shaped training step: backbone + 2xReNet(100) + MHA(2 heads, d_k=12, L=4096) + heads + "
                  "DiscriminativeLoss(C=24,K=32) + CE + Dice, fwd+bwd+clip+Adadelta, 256x256, batch 16 per GPU")


def inference_leg(model, dev, peaks, steps, warmup, rank=0, world=1, sampler=None, timed=None, with_cpu=True):
    """pred.py inference (configs[0]): value (inputs in HBM), e2e (host image in, host masks out), per-kernel times,
    bit-exactness flags against the oracle / scikit-learn, and (N = 1) the host-CPU reference beside it."""
    import torch
    from isa_b200 import _lib, synth
    from isa_b200.prediction import Prediction
    from isa_b200.settings import CVPPPModelSettings
    ms_ = CVPPPModelSettings()
    pred = Prediction(ms_.IMAGE_HEIGHT, ms_.IMAGE_WIDTH, ms_.MEAN, ms_.STD, False, model, 1, seed=0)
    raws = [synth.leaf_image(100 * rank + j, RAW_H, RAW_W) for j in range(4)]
    tens = [pred.image_to_tensor(r)[0].unsqueeze(0).to(dev) for r in raws]
    last = {}

    def step_dev(i):
        sem, emb = model.predict_device(tens[i % 4])
        last["o"] = pred.cluster_device(sem[0], emb[0], N_OBJ, RAW_H, RAW_W)

    def _raw_cycle():
        j = 0
        while True:
            yield raws[j % 4]
            j += 1

    e2e_results = pred.predict_many(_raw_cycle())    # the pipelined public path pred_list.py uses

    def step_e2e(i):
        last["masks"] = next(e2e_results)            # host image in -> host uint8 masks out

    if timed is None:
        def timed(step_fn, n_steps, n_warm):
            for i in range(n_warm):
                step_fn(i)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for i in range(n_steps):
                step_fn(n_warm + i)
            e.record()
            torch.cuda.synchronize()
            return s.elapsed_time(e)

    _lib.TIMER.reset()
    if sampler is not None:
        sampler.start()
    timed(step_dev, 0, warmup)
    _lib.TIMER.enabled = True
    ms = timed(step_dev, steps, 0)
    _lib.TIMER.enabled = False
    clocks = sampler.stop() if sampler is not None else None
    summ = _lib.TIMER.summary()
    launches = sum(_lib.KERNELS_PER_CALL[k] * c for k, (c, _) in summ.items())
    ms_e2e = timed(step_e2e, steps, 1)
    res = last["o"][4]
    n_pts = int(res.info[2])
    it_sum = int(res.n_iter.sum())
    km = summ.get("isa_kmeans_fit", (1, 0.0))
    per_launch_s = km[1] / max(km[0], 1) * 1e-3
    units = it_sum * n_pts * (4 * C_EMB + 4)
    ach = units / max(per_launch_s, 1e-12) / 1e9
    total_img = world * steps
    # parity flags against the oracle and the real scikit-learn on ALL benchmark images (outside the timed region)
    from oracle import kmeans as KM
    flags = {}
    if rank == 0:
        ident_o, ident_s, agree, inertia_d, self_agree = [], [], [], [], []
        for ti in range(len(tens)):
            sem, emb = model.predict_device(tens[ti])
            fg, X = KM.gather_foreground(sem[0].cpu().numpy(), emb[0].cpu().numpy())
            if len(X) < N_OBJ:
                continue
            o = KM.kmeans_oracle(X, N_OBJ, seed=0)
            got = pred.cluster_device(sem[0], emb[0], N_OBJ)[1].cpu().numpy()
            ident_o.append(bool(np.array_equal(got, KM.scatter_labels(fg, o["labels"]))))
            sk = KM.sklearn_fit_predict(X, N_OBJ, 0)
            ident_s.append(bool(KM.same_up_to_permutation(got[fg != 0], sk + 1)))
            agree.append(round(float(KM.partition_agreement(got[fg != 0], sk + 1)), 5))
            if not ident_s[-1]:
                # quantify the mismatch: scikit-learn's own fp32 sums depend on the BLAS kernel / thread count, and its
                # tol-based stop can end one Lloyd iteration earlier or later on near-tied inputs (random-init embeddings)
                def inertia(lab):
                    x = X.astype(np.float64)
                    return float(sum(((x[lab == c] - x[lab == c].mean(0)) ** 2).sum() for c in np.unique(lab)))
                inertia_d.append((inertia(got[fg != 0]) - inertia(sk + 1)) / inertia(sk + 1))
                try:
                    from threadpoolctl import threadpool_limits
                    with threadpool_limits(limits=1):
                        sk1 = KM.sklearn_fit_predict(X, N_OBJ, 0)
                    self_agree.append(round(float(KM.partition_agreement(sk1 + 1, sk + 1)), 5))
                except Exception:
                    pass
        flags["labels_identical_to_oracle"] = bool(ident_o) and all(ident_o)
        flags["labels_identical_to_sklearn_up_to_permutation"] = bool(ident_s) and all(ident_s)
        flags["images_checked"] = len(ident_o)
        flags["images_identical_to_sklearn"] = int(sum(ident_s))
        flags["sklearn_partition_agreement"] = min(agree) if agree else None
        if inertia_d:
            flags["inertia_rel_diff_vs_sklearn"] = inertia_d
        if self_agree:
            flags["sklearn_self_agreement_1_vs_all_threads"] = self_agree
    leg = {
        "value": total_img / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
        "config": {"workload": INFER_WORKLOAD,
                   "l2": "4 alternating images; per-image working set (35 restarts x labels + X) < L2 by design",
                   "fg_points": n_pts, "lloyd_restart_iterations": it_sum, "lloyd_grid_iterations": int(res.info[3]),
                   "kernel_ms_per_step": {k: round(v[1] / steps, 4) for k, v in sorted(summ.items())}, **flags},
        "e2e": {"value": total_img / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": 3 * NET_H * NET_W * 4,
                "d2h_bytes_per_step": 2 * RAW_H * RAW_W},
        "gpu_launches": launches,
        "roofline": {"kernel": "isa_kmeans_fit", "bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                     "frac": ach / peaks["hbm"], "traffic": ncu_traffic("isa_kmeans_fit"), "peak_source": peaks["src"],
                     "note": "X (n x 24 fp32) is L2 resident at this size; bytes = restart-iterations * n * (4C+4); "
                             "restart-iterations/s = %.0f" % (it_sum / max(per_launch_s, 1e-12)),
                     "avg_launch_us": per_launch_s * 1e6},
    }
    if clocks is not None:
        leg["clocks"] = clocks
    if with_cpu and rank == 0:
        leg["cpu_baseline"] = cpu_reference("infer", steps=1, warmup=0)
    return leg


def run_ours(args):
    import torch
    import torch.distributed as dist
    from isa_b200 import _lib, parallel
    from isa_b200.model import Model
    from isa_b200.prediction import Prediction
    from isa_b200.settings import CVPPPModelSettings, CVPPPTrainingSettings

    rank, world, local_rank = parallel.init_from_env("nccl")
    assert torch.cuda.is_available(), "bench.py needs a B200: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    peaks = measured_peaks()
    ts = CVPPPTrainingSettings()
    torch.manual_seed(ts.SEED)
    model = Model('CVPPP', 'ReSeg', ts.N_CLASSES, ts.MAX_N_OBJECTS, use_instance_segmentation=True,
                  n_embedding=C_EMB, distributed=world > 1, device=dev)
    sampler = ClockSampler(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup):
        for i in range(warmup):
            step_fn(i)
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            step_fn(warmup + i)
        e.record()
        barrier()
        return parallel.max_over_ranks(s.elapsed_time(e), dev)

    out = {}
    if args.workload == "train":
        model.define_criterion(ts.CLASS_WEIGHTS, ts.DELTA_VAR, ts.DELTA_DIST, ts.NORM, ts.OPTIMIZE_BG, ts.CRITERION)
        model.define_optimizer(ts.LEARNING_RATE, ts.WEIGHT_DECAY, ts.LR_DROP_FACTOR, ts.LR_DROP_PATIENCE, ts.OPTIMIZER)
        bs = PER_GPU_BATCH
        n_host = 2
        host = []
        for j in range(n_host):
            img, sem, ins, labels, nobj = train_batch(1000 * rank + j, bs)
            host.append([torch.from_numpy(a).pin_memory() for a in (img, sem, ins, nobj)])
        devb = [[t.to(dev) for t in hb] for hb in host]
        clip = ts.CLIP_GRAD_NORM
        last = {}

        def step_dev(i):
            b = devb[i % n_host]
            last["m"] = model.train_step(b[0], b[1], b[2], b[3], clip)

        from isa_b200.data import CudaPrefetcher

        class _Cycle(object):                        # an endless loader over the pinned host batches
            def __iter__(self):
                j = 0
                while True:
                    yield host[j % n_host]
                    j += 1

            def __len__(self):
                return 1 << 30

        e2e_batches = iter(CudaPrefetcher(_Cycle(), dev))   # the public staging path Model.fit uses (data.py)

        def step_e2e(i):
            b = next(e2e_batches)                    # H2D of this step's pinned inputs (overlaps the previous step)
            m = model.train_step(b[0], b[1], b[2], b[3], clip)
            last["loss"] = float(m['Cost'])          # device -> host read of the step's result

        # (a) per-kernel times: an EAGER pass with CUDA events around every C-ABI call (explains the step, is not the value)
        _lib.TIMER.reset()
        timed(step_dev, 0, args.warmup)              # warm-up outside the kernel timers
        _lib.TIMER.enabled = True
        ms_eager = timed(step_dev, args.steps, 0)
        _lib.TIMER.enabled = False
        summ = _lib.TIMER.summary()
        launches = sum(_lib.KERNELS_PER_CALL[k] * c for k, (c, _) in summ.items())
        # (b) value: inputs resident in HBM, the step replayed from its CUDA graph (Model.enable_cuda_graph: the same
        #     ~800 launches per step, submitted by one cudaGraphLaunch instead of one by one from Python)
        use_graph = not args.no_graph
        if use_graph:
            model.enable_cuda_graph(warmup_steps=1)
        sampler.start()
        ms = timed(step_dev, args.steps, 3)          # the first warm-up step here captures the graph
        clocks = sampler.stop()
        ms_e2e = timed(step_e2e, args.steps, 1)
        h2d = sum(t.numel() * t.element_size() for t in host[0])
        total_img = bs * world * args.steps
        tok = bs * (NET_H // 4) * (NET_W // 4)
        # split calls of one step: forward x (256, 200 x3) + backward [dgx|dghn] (800) and [x|h|h] (cin+200) per sweep
        split_elems = tok * ((256 + 3 * 200) + 4 * 800 + (256 + 3 * 200) + 4 * 200)
        n_split = max(summ.get("isa_split_bf16x3", (1, 0))[0] // args.steps, 1)
        cfg = {"bs": bs, "tgt_bytes": 8 * K_MAX, "split_bytes_per_call": split_elems * (4 + 6) / n_split}
        # dominant kernel among ours, and the same figures for every timed entry point
        dom = max(summ.items(), key=lambda kv: kv[1][1])
        dom_roof = roofline_entry(dom[0], dom[1][0], dom[1][1], cfg, peaks)
        all_roof = [r for r in (roofline_entry(k_, v_[0], v_[1], cfg, peaks) for k_, v_ in sorted(summ.items())) if r]
        kernel_ms = {k: round(v[1] / args.steps, 4) for k, v in sorted(summ.items())}
        out = {
            "metric": METRIC,
            "value": total_img / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": TRAIN_WORKLOAD,
                       "per_gpu_batch": bs, "global_batch": bs * world, "parallelism": "dp%d" % world,
                       "l2": "working set (activations) >> 126 MB L2, two alternating input batches",
                       "target_format_value": "int64 one-hot (reference collate)", "kernel_ms_per_step": kernel_ms,
                       "cuda_graph": bool(use_graph), "ms_per_step_eager_with_kernel_timers": ms_eager / args.steps},
            "clocks": clocks,
            "e2e": {"value": total_img / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "roofline": dom_roof,
            "rooflines": all_roof,
        }
        if world == 1 and not args.no_inference:
            # pred.py runs a checkpoint, not the net this bench has just trained for a few steps: a fresh model with
            # the settings' seed, i.e. exactly what `--workload infer` measures
            del model
            torch.cuda.empty_cache()
            torch.manual_seed(ts.SEED)
            infer_model = Model('CVPPP', 'ReSeg', ts.N_CLASSES, ts.MAX_N_OBJECTS, use_instance_segmentation=True,
                                n_embedding=C_EMB, device=dev)
            out["inference"] = inference_leg(infer_model, dev, peaks, steps=5, warmup=3)
    else:
        leg = inference_leg(model, dev, peaks, steps=args.steps, warmup=args.warmup, rank=rank, world=world, sampler=sampler,
                            timed=timed, with_cpu=False)
        out = {
            "metric": METRIC,
            "value": leg["value"], "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": leg["config"], "clocks": leg["clocks"], "e2e": leg["e2e"], "gpu_launches": leg["gpu_launches"],
            "roofline": leg["roofline"],
        }
    if rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_reference(args.workload, steps=1, warmup=0)
    if rank == 0:
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference(workload, steps=1, warmup=0):
    """Times the reference's CPU implementation of the workload on a bounded sample; returns the cpu_baseline dict."""
    import torch
    from oracle import kmeans as KM
    from oracle.model_ref import ReSegRef, discriminative_loss_torch
    from isa_b200 import synth
    from isa_b200.settings import CVPPPTrainingSettings
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ts = CVPPPTrainingSettings()
    torch.manual_seed(ts.SEED)
    net = ReSegRef(ts.N_CLASSES, n_embedding=C_EMB)
    if workload == "train":
        bs = 2
        img, sem, ins, labels, nobj = train_batch(0, bs)
        img, sem, ins = torch.from_numpy(img), torch.from_numpy(sem), torch.from_numpy(ins).float()
        opt = torch.optim.Adadelta(net.parameters(), lr=ts.LEARNING_RATE, weight_decay=ts.WEIGHT_DECAY)
        ce = torch.nn.CrossEntropyLoss()

        def step():
            net.train()
            sem_out, emb = net(True, img)
            ins_cost, _ = discriminative_loss_torch(emb, ins, nobj, K_MAX, ts.DELTA_VAR, ts.DELTA_DIST, ts.NORM)
            probs = torch.softmax(sem_out, 1)
            semf = sem.float()
            dice = (2 * (probs * semf).sum((2, 3)) + 1.0) / (probs.sum((2, 3)) + semf.sum((2, 3)) + 1.0)
            cost = ins_cost + ce(sem_out, sem.max(1)[1]) + (1 - dice[:, 1:].mean(1)).mean()
            opt.zero_grad()
            cost.backward()
            torch.nn.utils.clip_grad_norm_(net.parameters(), ts.CLIP_GRAD_NORM)
            opt.step()

        for _ in range(warmup):
            step()
        t0 = time.time()
        for _ in range(steps):
            step()
        dt = (time.time() - t0) / steps
        return {"value": bs / dt, "unit": "images/s", "cores": cores, "kind": "port",
                "sample": "batch %d of the %d-image step (same graph from reference ops on the CPU: nn.GRU ReNet, reference MHA math, "
                          "the reference's broadcast discriminative-loss graph + autograd, Adadelta), %d step(s)" % (bs, PER_GPU_BATCH, steps),
                "seconds_per_step": dt}
    net.eval()
    raw = synth.leaf_image(0, RAW_H, RAW_W)
    from PIL import Image
    from isa_b200.settings import CVPPPModelSettings
    ms_ = CVPPPModelSettings()

    def step():
        img = Image.fromarray(raw).resize((NET_W, NET_H), Image.BILINEAR)
        x = (np.asarray(img, dtype=np.float32).transpose(2, 0, 1) / 255.0 - np.asarray(ms_.MEAN, np.float32).reshape(3, 1, 1)) / np.asarray(ms_.STD, np.float32).reshape(3, 1, 1)
        with torch.no_grad():
            sem_out, emb = net(False, torch.from_numpy(x).unsqueeze(0))
            sem_p = torch.softmax(sem_out, 1)
        fg, mask = KM.cluster_reference(sem_p[0].numpy(), emb[0].numpy(), N_OBJ, seed=0, impl="sklearn")
        return KM.upsample_nearest(fg, RAW_H, RAW_W), KM.upsample_nearest(mask, RAW_H, RAW_W)

    for _ in range(warmup):
        step()
    t0 = time.time()
    for _ in range(steps):
        step()
    dt = (time.time() - t0) / steps
    return {"value": 1.0 / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "%d whole image(s): CPU net forward (reference ops) + real scikit-learn KMeans(k=16, n_init=35, max_iter=500) "
                      "+ numpy scatter + cv2 INTER_NEAREST" % steps,
            "seconds_per_step": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference(args.workload, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1))
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": cb["value"], "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cb["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": TRAIN_WORKLOAD if args.workload == "train" else INFER_WORKLOAD,
                   "arm": "host-CPU reference implementation of this workload, bounded sample (see cpu_baseline.sample)"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = None


def _quarantine_stdout():
    """Libraries print to fd 1 (NCCL's version banner, cuDNN notes): point fd 1 at stderr for the run and keep the
    real stdout for the ONE JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "infer"])
    ap.add_argument("--no-graph", dest="no_graph", action="store_true", help="keep the training step eager (no CUDA-graph replay)")
    ap.add_argument("--no-inference", dest="no_inference", action="store_true",
                    help="train workload: skip the extra pred.py inference leg reported under \"inference\" at N = 1")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    _quarantine_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
