#!/usr/bin/env python
"""bench.py — AutoFormer-S supernet random-sample training step on the B200 engine.

    python bench.py --gpus N --steps K --warmup W            # own arm (B200 kernels)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port)

Metric (BASELINE.json): supernet images/sec, 224^2, batch 128 per GPU.  One "step" = one pass
of the hot path over one synthetic batch: sample a subnet (supernet_engine.py:13-24, same RNG
stream on every rank), forward, cross-entropy, backward, per-layer gradient all-reduce,
AdamW step.  `value` is timed with the batch already resident in HBM; `e2e` is the same step
through the public API (SupernetTrainer.step) with the batch coming from pinned host memory
and the loss read back to the host every step.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "supernet images/sec (224^2, bs128/GPU)"
UNIT = "images/s"
PER_GPU_BATCH = 128
WORKLOAD = ("AutoFormer-S supernet random-sample training step (embed 320-448, heads 5-7, depth 12-14, "
            "mlp 3-4), 224^2, bs128/GPU, relative position on K and V, DropPath 0.1, AdamW")
# BASELINE.json configs that are bench lines (the others are parity-test cases): c3 is the headline
# (`metric` is quoted on it), c5 the largest search space.
CONFIGS = {
    "c2": dict(size="deit_s", batch=256, metric="DeiT-S + iRPE training images/sec (224^2, bs256)",
               workload=("DeiT-S + iRPE (product method, contextual mode on keys, 50 buckets, shared head) training step, "
                         "224^2, bs256, 12 blocks, DropPath 0.1, AdamW; 1 GPU")),
    "c3": dict(size="S", batch=128, metric=METRIC, workload=WORKLOAD),
    "c4": dict(size="clip_b32", batch=128, metric="TinyCLIP ViT-B/32 contrastive training image-text pairs/sec (bs128/GPU)",
               workload=("CLIP ViT-B/32 (image tower 12 x 768, 32x32 patches, 50 tokens; text tower 12 x 512, 77 tokens, "
                         "causal mask) contrastive training step, 224^2, bs128/GPU (1024 on 8 GPUs), local loss + feature "
                         "all_gather with gradient, AdamW")),
    "c5": dict(size="B", batch=64, metric="supernet images/sec (224^2, bs64/GPU)",
               workload=("AutoFormer-B supernet random-sample training step (embed 528-624, heads 9-10, depth 14-16, "
                         "mlp 3-4), 224^2, bs64/GPU, relative position on K and V, DropPath 0.1, AdamW")),
}


def flops_per_image(cfg, n_tokens=197):
    """Algorithmic forward FLOPs per image of a sampled subnet (SURVEY.md §8d formulas)."""
    E = cfg["embed_dim"][0]
    N = n_tokens
    total = 2 * 196 * 768 * E
    attn = 0
    for i in range(cfg["layer_num"]):
        h = cfg["num_heads"][i]
        ffn = int(E * cfg["mlp_ratio"][i])
        total += 2 * N * (E * 3 * 64 * h + 64 * h * E + 2 * E * ffn)
        attn += 4 * h * N * N * 64 + 2 * h * N * 64 * (60 + 60)
    total += 2 * E * 1000
    return total + attn, attn


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_baseline(steps: int, warmup: int, batch: int = 4, config: str = "c3"):
    """The reference's CPU path for this workload on the host cores: the reference's OWN
    Vision_TransformerSuper (unmodified, loaded from the staged checkout: kind "reference") when
    available, else the oracle port (PyTorch fp32 restatement: kind "port"); forward + autograd
    backward + AdamW on a bounded sample (small batch, same config stream)."""
    import torch
    import torch.nn.functional as F
    from oracle import refload, vit_oracle as vo
    size = CONFIGS[config]["size"]
    if size == "deit_s":
        return cpu_baseline_deit(steps, warmup, batch)
    if size == "clip_b32":
        return cpu_baseline_clip(steps, warmup, max(batch, 8))
    spec = {"S": vo.SUPERNET_S, "B": vo.SUPERNET_B, "T": vo.SUPERNET_T}[size]
    space = vo.SEARCH_SPACE[size]
    sd0 = vo.init_params(spec, seed=0)
    if refload.available():
        kind = "reference"
        net = refload.autoformer("reference").Vision_TransformerSuper(
            img_size=224, patch_size=16, embed_dim=spec.embed_dim, depth=spec.depth, num_heads=spec.num_heads,
            mlp_ratio=spec.mlp_ratio, qkv_bias=True, drop_rate=0.0, drop_path_rate=0.1, gp=True, num_classes=1000,
            max_relative_position=14, relative_position=True, change_qkv=True, abs_pos=True)
        net.load_state_dict(sd0)
        net.train()
        params = list(net.parameters())

        def fwd(cfg, images):
            net.set_sample_config(cfg)
            return net(images)
        what = "AutoFormer/model/supernet_transformer.py (unmodified, staged checkout)"
    else:
        kind = "port"
        sd = {k: v.requires_grad_(True) for k, v in sd0.items()}
        params = list(sd.values())
        fwd = lambda cfg, images: vo.supernet_forward(sd, cfg, images, spec)
        what = "oracle/vit_oracle.py"
    opt = torch.optim.AdamW(params, lr=5e-4, weight_decay=0.05)
    rnd = random.Random(0)
    torch.manual_seed(0)
    images = torch.randn(batch, 3, 224, 224)
    targets = torch.randint(0, 1000, (batch,))
    # give the CPU path its best thread count (oversubscribing a 128-thread host with a batch of
    # 4 is ~50x slower than 16-32 threads): one probe step per candidate, keep the fastest
    ncpu = os.cpu_count() or 1
    probe_cfg = vo.sample_configs(space, random.Random(1))
    best = (float("inf"), ncpu)
    for n in sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(n)
        t0 = time.perf_counter()
        F.cross_entropy(fwd(probe_cfg, images), targets).backward()
        dt = time.perf_counter() - t0
        if dt < best[0]:
            best = (dt, n)
        if dt > 20.0:
            break
    torch.set_num_threads(best[1])
    opt.zero_grad(set_to_none=True)
    times = []
    for s in range(warmup + steps):
        cfg = vo.sample_configs(space, rnd)
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = F.cross_entropy(fwd(cfg, images), targets)
        loss.backward()
        opt.step()
        loss.item()
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": batch * len(times) / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{len(times)} training steps of batch {batch} (same supernet-{size} config stream), fp32, "
                      f"{what} on {torch.get_num_threads()} host threads"}, total / len(times)


def _e2e_steps(tr, host_a, host_b, n):
    """n end-to-end steps through trainer.stage / trainer.step: every step's batch is copied from pinned host memory
    inside the timed region (side stream, one step ahead) and every step's loss is copied to pinned host memory and
    read on the host one step behind the enqueue front.  Returns (host ms per step list, last loss)."""
    import torch
    pins = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    host, pending, last = [], None, 0.0
    staged = tr.stage(host_a[0], host_b[0])
    for s in range(n):
        t0 = time.perf_counter()
        loss = tr.step(staged)
        buf = pins[s & 1]
        buf.copy_(loss, non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        if s + 1 < n:
            staged = tr.stage(host_a[(s + 1) % len(host_a)], host_b[(s + 1) % len(host_b)])
        if pending is not None:
            pending[1].synchronize()
            last = float(pending[0])
        pending = (buf, done)
        if s + 1 == n:
            done.synchronize()
            last = float(buf)
        host.append((time.perf_counter() - t0) * 1e3)
    return host, last


def gpu_reference_supernet(spec, cfg_warm, cfg_timed, dev_imgs, dev_tgts, B):
    """The reference's own Vision_TransformerSuper (unmodified files, stock PyTorch ops) training on the SAME
    GPU, batch and subnet stream: fp16 autocast + loss scaling (supernet_engine.py:65-84) + fused AdamW."""
    import torch
    import torch.nn.functional as F
    from oracle import refload
    if not refload.available():
        return None
    try:
        dev = dev_imgs[0].device
        torch.manual_seed(0)
        net = refload.autoformer("reference").Vision_TransformerSuper(
            img_size=224, patch_size=16, embed_dim=spec["embed_dim"], depth=spec["depth"], num_heads=spec["num_heads"],
            mlp_ratio=spec["mlp_ratio"], qkv_bias=True, drop_rate=0.0, drop_path_rate=0.1, gp=True, num_classes=1000,
            max_relative_position=14, relative_position=True, change_qkv=True, abs_pos=True).to(dev).train()
        opt = torch.optim.AdamW(net.parameters(), lr=5e-4, weight_decay=0.05, fused=True)
        scaler = torch.amp.GradScaler("cuda")

        def step(cfg, x, y):
            opt.zero_grad(set_to_none=True)
            net.set_sample_config(cfg)
            with torch.autocast("cuda", dtype=torch.float16):
                loss = F.cross_entropy(net(x), y)
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
        n = len(dev_imgs)
        for s, cfg in enumerate(cfg_warm):
            step(cfg, dev_imgs[s % n], dev_tgts[s % n])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s, cfg in enumerate(cfg_timed):
            step(cfg, dev_imgs[s % n], dev_tgts[s % n])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out = {"value": B * len(cfg_timed) / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / len(cfg_timed),
               "what": "reference Vision_TransformerSuper (AutoFormer/model/supernet_transformer.py + module/*, unmodified; cuBLAS / "
                       "stock PyTorch attention with the relative-position einsums) under fp16 autocast + GradScaler + fused AdamW, "
                       f"same GPU, same batch, the first {len(cfg_timed)} subnets of the timed stream"}
        del net, opt
        torch.cuda.empty_cache()
        return out
    except Exception as e:   # noqa: BLE001
        return {"unavailable": repr(e)[:300]}


def _reference_deit(device):
    """The reference's own DeiT-S + iRPE VisionTransformer (unmodified files from the staged checkout)."""
    import torch
    from functools import partial
    from oracle import refload
    over = "reference"
    if device != "cpu":
        over = "reference_cuda" if refload.reference_rpe_ops_cuda() is not None else "reference"
    vit = refload.rpe_vision_transformer(over)
    cfg = vit.irpe.get_rpe_config(ratio=1.9, method='product', mode='ctx', shared_head=True, skip=1, rpe_on='k')
    torch.manual_seed(0)
    net = vit.VisionTransformer(patch_size=16, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, qkv_bias=True,
                                drop_path_rate=0.1, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), rpe_config=cfg)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if "lookup_table" in n:
                p.normal_(std=0.02)
    return net.to(device).train(), over


def cpu_baseline_deit(steps: int, warmup: int, batch: int = 4):
    import torch
    import torch.nn.functional as F
    from oracle import refload
    assert refload.available(), "config c2's CPU arm needs the staged reference (scripts/stage_reference.py)"
    net, _ = _reference_deit("cpu")
    opt = torch.optim.AdamW(net.parameters(), lr=5e-4, weight_decay=0.05)
    torch.manual_seed(0)
    images, targets = torch.randn(batch, 3, 224, 224), torch.randint(0, 1000, (batch,))
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = F.cross_entropy(net(images), targets)
        loss.backward()
        opt.step()
        loss.item()
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": batch * len(times) / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "reference",
            "sample": f"{len(times)} training steps of batch {batch}, fp32, iRPE/DeiT-with-iRPE rpe_vision_transformer.py "
                      f"(unmodified, pure-PyTorch rpe fallback) on {torch.get_num_threads()} host threads"}, total / len(times)


def bench_deit(args):
    """BASELINE config 2: DeiT-S + iRPE (product, contextual, keys) bs256 on ONE B200: the native
    runtime's training step against the reference model + the reference rpe_ops CUDA build on the same GPU."""
    import torch
    import torch.nn.functional as F
    from cream_b200 import _lib, ops
    from cream_b200.deit import DeitIrpe, DeitTrainer
    from oracle import refload
    assert int(os.environ.get("WORLD_SIZE", "1")) == 1, "config c2 is a single-GPU configuration"
    conf = CONFIGS["c2"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _lib.load()
    W, K, B = max(3, args.warmup), args.steps, args.batch or conf["batch"]
    torch.manual_seed(0)
    net = DeitIrpe(drop_path_rate=0.1).to(dev).train()
    tr = DeitTrainer(net)
    g = torch.Generator().manual_seed(1234)
    n_host = 4
    host_imgs = [torch.randn(B, 3, 224, 224, generator=g).pin_memory() for _ in range(n_host)]
    host_tgts = [torch.randint(0, 1000, (B,), generator=g).pin_memory() for _ in range(n_host)]
    dev_imgs, dev_tgts = [t.to(dev) for t in host_imgs], [t.to(dev) for t in host_tgts]
    pin_loss = torch.zeros((), dtype=torch.float32).pin_memory()

    def timed(n, from_host):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.LAUNCHES[0]
        host = []
        e0.record()
        last = 0.0
        last = 0.0
        if from_host:
            host, last = _e2e_steps(tr, host_imgs, host_tgts, n)
        else:
            for s in range(n):
                t0 = time.perf_counter()
                tr.step(dev_imgs[s % n_host], dev_tgts[s % n_host])
                host.append((time.perf_counter() - t0) * 1e3)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), _lib.LAUNCHES[0] - l0, host, last

    timed(W, False)
    sampler = ClockSampler(0)
    sampler.start()
    ms, launches, host, _ = timed(K, False)
    timed(2, True)
    ms_e2e, _, _, last_loss = timed(K, True)
    clocks = sampler.stop()

    # roofline of the kernel north_star names (fused attention + iRPE): the attention kernels of one block
    # at this configuration's shape, timed alone with CUDA events after the step measurements
    H, N = 6, 197
    qkv = ops.empty_bf16(B * N, 3 * 64 * H)
    qkv.copy_(torch.randn(B * N, 3 * 64 * H, device=dev))
    P = dict(net.named_parameters())
    r = tr.native
    tk = ops.new_pack(1, dev)
    t0 = P["blocks.0.attn.rpe_k.lookup_table_weight"].detach()
    ops.pack_tables(tk, 1, t0, t0.shape[2], 0, (t0.stride(0), t0.stride(2), t0.stride(1)))
    from cream_b200 import ops as _o
    ids, nb = _o.irpe_bucket_ids(3, 14, 14, 1, 1.9, 3.8, 15.2)
    it = _o.irpe_index_table_u8(ids, dev)
    gp = (14,) + tuple(_o.irpe_grid_product_structure(ids, 14, 1))     # the structured gather the model runs
    dout = ops.empty_bf16(B * N, 64 * H)
    dout.copy_(torch.randn(B * N, 64 * H, device=dev))
    def attn_times(reps=10):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        out, lse = ops.attention_fwd(qkv, B, H, N, 0.125, tk=tk, idx=(it, None, None, None), gp=gp)
        ops.attention_bwd(qkv, out, lse, dout, B, H, N, 0.125, tk=tk, idx=(it, None, None, None), gp=gp)
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(reps):
            out, lse = ops.attention_fwd(qkv, B, H, N, 0.125, tk=tk, idx=(it, None, None, None), gp=gp)
        ev[1].record()
        for _ in range(reps):
            ops.attention_bwd(qkv, out, lse, dout, B, H, N, 0.125, tk=tk, idx=(it, None, None, None), gp=gp)
        ev[2].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / reps * 1e-3, ev[1].elapsed_time(ev[2]) / reps * 1e-3
    t_fwd, t_bwd = attn_times()
    fl_fwd = 4.0 * B * H * N * N * 64 + 2.0 * B * H * N * 64 * 64
    by_fwd = 4.0 * B * H * N * 64 * 2
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm, peak_tf = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops_sustained", 1400.0)
    roofline = {"kernel": "attn_fwd_kernel, iRPE contextual product gather on keys, grid-product structured path (the fused attention+iRPE kernel as the model runs it)",
                "bound": "hbm", "achieved": by_fwd / t_fwd / 1e9, "peak": hbm, "unit": "GB/s", "frac": by_fwd / t_fwd / 1e9 / hbm,
                "tflops": fl_fwd / t_fwd / 1e12, "tflops_frac_of_peak": fl_fwd / t_fwd / 1e12 / peak_tf, "avg_launch_us": t_fwd * 1e6,
                "bwd_avg_us": t_bwd * 1e6, "bwd_tflops": 2.5 * fl_fwd / t_bwd / 1e12, "traffic": None,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"}

    # ---- GPU-side reference on the same B200: reference model + reference rpe_ops CUDA build ----
    gpu_ref = None
    if refload.available():
        try:
            ref, over = _reference_deit(dev)
            opt = torch.optim.AdamW(ref.parameters(), lr=5e-4, weight_decay=0.05, fused=True)
            scaler = torch.amp.GradScaler("cuda")
            def ref_step(x, y):
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.float16):     # the reference trains with fp16 autocast (engine.py)
                    loss = F.cross_entropy(ref(x), y)
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()
                return loss
            for s in range(3):
                ref_step(dev_imgs[s % n_host], dev_tgts[s % n_host])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            kk = max(3, K // 2)
            for s in range(kk):
                ref_step(dev_imgs[s % n_host], dev_tgts[s % n_host])
            e1.record()
            torch.cuda.synchronize()
            rms = e0.elapsed_time(e1)
            gpu_ref = {"value": B * kk / (rms * 1e-3), "unit": UNIT, "ms_per_step": rms / kk,
                       "what": f"reference VisionTransformer (rpe_vision_transformer.py, unmodified) + rpe_ops={over}, "
                               "fp16 autocast + GradScaler + fused AdamW, same GPU, same batch"}
            del ref, opt
        except Exception as e:   # noqa: BLE001
            gpu_ref = {"unavailable": repr(e)[:300]}
    imgs = B * K
    line = {"metric": conf["metric"], "value": imgs / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": conf["workload"], "baseline_config": "c2", "global_batch": B, "per_gpu_batch": B,
                       "parallelism": "dp1", "l2": "activations per step (> 5 GB) exceed the 126 MB L2; 4 rotating batches"},
            "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * 3 * 224 * 224 * 4 + B * 8,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K, "last_loss": last_loss},
            "gpu_launches": launches, "host_enqueue_ms_per_step": {"mean": sum(host) / len(host), "min": min(host)},
            "clocks": clocks, "roofline": roofline, "gpu_reference": gpu_ref}
    if gpu_ref and "value" in gpu_ref:
        line["speedup_vs_gpu_reference"] = line["value"] / gpu_ref["value"]
    if not args.no_cpu_baseline and refload.available():
        line["cpu_baseline"], _ = cpu_baseline_deit(1, 1)
    print(json.dumps(line))


def _clip_batches(B, n, seed, pin=True):
    import torch
    g = torch.Generator().manual_seed(seed)
    pinned = (lambda t: t.pin_memory()) if pin else (lambda t: t)
    imgs = [pinned(torch.randn(B, 3, 224, 224, generator=g)) for _ in range(n)]
    txts = []
    for _ in range(n):
        t = torch.randint(1, 49406, (B, 77), generator=g)
        eot = torch.randint(4, 77, (B,), generator=g)
        for b in range(B):
            t[b, eot[b]] = 49407
            t[b, eot[b] + 1:] = 0
        txts.append(pinned(t))
    return imgs, txts


def _reference_clip(device):
    """The reference's own CLIP ViT-B/32 (TinyCLIP/src/open_clip/model.py + loss.py, unmodified)."""
    import torch
    from cream_b200.clip import VIT_B_32 as C4
    from oracle import refload
    m = refload.open_clip_model()
    torch.manual_seed(0)
    net = m.CLIP(C4["embed_dim"], dict(C4["vision_cfg"]), dict(C4["text_cfg"])).to(device).train()
    named = list(net.named_parameters())
    skip = lambda n, p: p.ndim < 2 or "bn" in n or "ln" in n or "bias" in n or "logit_scale" in n
    opt = torch.optim.AdamW([dict(params=[p for n, p in named if skip(n, p)], weight_decay=0.0),
                             dict(params=[p for n, p in named if not skip(n, p)], weight_decay=0.2)],
                            lr=5e-4, betas=(0.9, 0.98), eps=1e-6, fused=(device != "cpu"))
    return net, opt, refload.open_clip_loss().ClipLoss()


def cpu_baseline_clip(steps: int, warmup: int, batch: int = 8):
    import torch
    from oracle import refload
    assert refload.available(), "config c4's CPU arm needs the staged reference (scripts/stage_reference.py)"
    net, opt, loss_fn = _reference_clip("cpu")
    imgs, txts = _clip_batches(batch, 1, 7, pin=False)
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        fi, ft, sc = net(imgs[0], txts[0])
        loss = loss_fn(fi, ft, sc)
        loss.backward()
        opt.step()
        loss.item()
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": batch * len(times) / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "reference",
            "sample": f"{len(times)} contrastive training steps of batch {batch}, fp32, TinyCLIP/src/open_clip model.py + "
                      f"loss.py (unmodified) on {torch.get_num_threads()} host threads"}, total / len(times)


def bench_clip(args):
    """BASELINE config 4: CLIP ViT-B/32 contrastive training, bs128/GPU, one process per GPU; the features
    are exchanged with all_gather (the path's one real collective besides the gradient average)."""
    import torch
    import torch.distributed as dist
    from cream_b200 import _lib, clip
    from oracle import refload
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    conf = CONFIGS["c4"]
    W, K, B = max(3, args.warmup), args.steps, args.batch or conf["batch"]
    torch.manual_seed(0)
    net = clip.CLIP(clip.VIT_B_32["embed_dim"], clip.VIT_B_32["vision_cfg"], clip.VIT_B_32["text_cfg"]).to(dev).train()
    tr = clip.ClipTrainer(net, graph=(world == 1 and not args.no_graph))   # fixed configuration: one captured CUDA graph per step
    n_host = 4
    host_imgs, host_txts = _clip_batches(B, n_host, 1234 + rank)
    dev_imgs, dev_txts = [t.to(dev) for t in host_imgs], [t.to(dev) for t in host_txts]
    pin_loss = torch.zeros((), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, from_host):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.LAUNCHES[0]
        host = []
        e0.record()
        last = 0.0
        if from_host:
            host, last = _e2e_steps(tr, host_imgs, host_txts, n)
        else:
            for s in range(n):
                t0 = time.perf_counter()
                tr.step(dev_imgs[s % n_host], dev_txts[s % n_host])
                host.append((time.perf_counter() - t0) * 1e3)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), _lib.LAUNCHES[0] - l0, host, last

    timed(W, False)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches, host, _ = timed(K, False)
    timed(2, True)
    ms_e2e, _, _, last_loss = timed(K, True)
    clocks = sampler.stop() if sampler else None

    # per-kernel-class times of one step (python sequencing -> ops.PROFILE brackets every launch)
    from cream_b200 import ops
    ops.PROFILE = []
    tr._step_eager(dev_imgs[0], dev_txts[0])      # eager: every launch bracketed (the timed regions replay the graph)
    torch.cuda.synchronize()
    by_kind = {}
    for kind, a, b, fl, by in ops.PROFILE:
        d = by_kind.setdefault(kind, [0.0, 0, 0.0, 0.0])
        d[0] += a.elapsed_time(b)
        d[1] += 1
        d[2] += fl
        d[3] += by
    ops.PROFILE = None
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm, peak_tf = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops_sustained", 1400.0)
    gm = by_kind.get("gemm", [1.0, 0, 0.0, 0.0])
    roofline = {"kernel": "gemm_bf16_kernel (tcgen05; all GEMMs of one step, event-bracketed per launch)", "bound": "tensor",
                "achieved": gm[2] / (gm[0] * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": gm[2] / (gm[0] * 1e-3) / 1e12 / peak_tf, "traffic": None,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                "per_kind_ms": {k: {"ms": round(v[0], 3), "launches": v[1]} for k, v in sorted(by_kind.items())}}

    gpu_ref = None
    if world == 1 and refload.available():
        try:
            ref, opt, loss_fn = _reference_clip(dev)
            def ref_step(x, t):
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16):      # --precision amp_bfloat16
                    fi, ft, sc = ref(x, t)
                    loss = loss_fn(fi, ft, sc)
                loss.backward()
                opt.step()
                return loss
            for s in range(3):
                ref_step(dev_imgs[s % n_host], dev_txts[s % n_host])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            kk = max(3, K // 2)
            for s in range(kk):
                ref_step(dev_imgs[s % n_host], dev_txts[s % n_host])
            e1.record()
            torch.cuda.synchronize()
            rms = e0.elapsed_time(e1)
            gpu_ref = {"value": B * kk / (rms * 1e-3), "unit": UNIT, "ms_per_step": rms / kk,
                       "what": "reference CLIP (open_clip/model.py + loss.py, unmodified: nn.MultiheadAttention, cuDNN conv, "
                               "cuBLAS) under bf16 autocast + fused AdamW, same GPU, same batch"}
            del ref, opt
        except Exception as e:   # noqa: BLE001
            gpu_ref = {"unavailable": repr(e)[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pairs = B * K * world
    line = {"metric": conf["metric"], "value": pairs / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": conf["workload"], "baseline_config": "c4", "global_batch": B * world, "per_gpu_batch": B,
                       "parallelism": f"dp{world}", "l2": "activations per step (> 3 GB) exceed the 126 MB L2; 4 rotating batches",
                       "engine": ("one CUDA graph per step (captured from the python sequencing of the C-ABI launches, replayed)"
                                  if getattr(tr, "_graph", None) is not None else
                                  "python sequencing of the C-ABI launches (one autograd node per tower)")},
            "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": world * (B * 3 * 224 * 224 * 4 + B * 77 * 8), "d2h_bytes_per_step": 4 * world,
                    "ms_per_step": ms_e2e / K, "last_loss": last_loss},
            "gpu_launches": launches, "host_enqueue_ms_per_step": {"mean": sum(host) / len(host), "min": min(host)},
            "clocks": clocks, "roofline": roofline, "gpu_reference": gpu_ref}
    if gpu_ref and "value" in gpu_ref:
        line["speedup_vs_gpu_reference"] = line["value"] / gpu_ref["value"]
    if not args.no_cpu_baseline and world == 1 and refload.available():
        line["cpu_baseline"], _ = cpu_baseline_clip(1, 1)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    conf = CONFIGS[args.config]
    base, sec_per_step = cpu_baseline(args.steps, max(1, min(args.warmup, 2)), config=args.config)
    line = {"impl": "reference", "metric": conf["metric"], "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": conf["workload"], "note": f"reference CPU path ({base['kind']}), bounded sample"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cream_b200", choices=["cream_b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS), help="BASELINE.json configuration")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default = the configuration's own)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap", type=int, default=1,
                    help="gradient all-reduce collectives per step at N > 1: 1 = one after the backward, k = k-1 "
                         "overlapped + one after, 0 = one per layer (round-1 schedule)")
    ap.add_argument("--python-engine", action="store_true", help="sequence the kernels from Python (cross-check)")
    ap.add_argument("--no-graph", action="store_true", help="c4: run the step eagerly instead of replaying its CUDA graph")
    ap.add_argument("--quick", action="store_true",
                    help="A/B runs (c3 / c5): the two timed regions only - no per-kernel pass, no reference legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "c2":
        return bench_deit(args)
    if args.config == "c4":
        return bench_clip(args)

    import torch
    import torch.distributed as dist
    from cream_b200 import _lib, ops
    from cream_b200.autoformer.model.supernet_transformer import Vision_TransformerSuper
    from cream_b200.trainer import SupernetTrainer, sample_configs
    from cream_b200.configs import SEARCH_SPACE, SUPERNETS

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    conf = CONFIGS[args.config]
    W = max(3, args.warmup)
    K = args.steps
    B = args.batch or conf["batch"]

    spec = SUPERNETS[conf["size"]]
    ss = SEARCH_SPACE[conf["size"]]
    torch.manual_seed(0)
    model = Vision_TransformerSuper(img_size=224, patch_size=16, embed_dim=spec["embed_dim"], depth=spec["depth"],
                                    num_heads=spec["num_heads"], mlp_ratio=spec["mlp_ratio"], qkv_bias=True, drop_rate=0.0,
                                    drop_path_rate=0.1, gp=True, num_classes=1000, max_relative_position=14,
                                    relative_position=True, change_qkv=True, abs_pos=True).to(dev).train()
    trainer = SupernetTrainer(model, ss, native=not args.python_engine, overlap=(True if args.overlap == 0 else args.overlap))
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 4
    host_imgs = [torch.randn(B, 3, 224, 224, generator=g).pin_memory() for _ in range(n_host)]
    host_tgts = [torch.randint(0, 1000, (B,), generator=g).pin_memory() for _ in range(n_host)]
    dev_imgs = [t.to(dev) for t in host_imgs]     # 77 MB each: > 126 MB L2 in aggregate with activations
    dev_tgts = [t.to(dev) for t in host_tgts]

    # ONE list of subnet configurations (sample_configs with random.Random(0): identical on every rank,
    # supernet_engine.py:36) shared by the warm-up, the `value` region, the `e2e` region and the
    # per-kernel roofline pass, so that all of them time the same work.
    rnd = random.Random(0)
    cfg_warm = [sample_configs(ss, rnd) for _ in range(W)]
    cfg_timed = [sample_configs(ss, rnd) for _ in range(K)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]
    loss_pins = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]

    def timed(cfgs, from_host):
        n_steps = len(cfgs)
        flops = attn_flops = 0.0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = _lib.LAUNCHES[0]
        e0.record()
        loss_host = 0.0
        per_step = []
        # end to end: every step's batch is copied from pinned host memory inside the timed region;
        # the copy of step s+1 is started (side stream) before the host waits for the loss of step s
        # and every step's loss is copied to pinned host memory and read on the host, one step
        # behind the enqueue front (the host reads loss s-1 while the GPU runs step s), so a host
        # that is faster than the GPU never drains the launch queue.
        staged = trainer.stage(host_imgs[0], host_tgts[0]) if from_host else None
        pending = None
        for s, cfg in enumerate(cfgs):
            t_host0 = time.perf_counter()
            i = s % n_host
            if from_host:
                loss = trainer.step(staged, config=cfg)
                buf = loss_pins[s & 1]
                buf.copy_(loss, non_blocking=True)          # device -> pinned host, every step
                done = torch.cuda.Event()
                done.record()
                if s + 1 < n_steps:
                    staged = trainer.stage(host_imgs[(s + 1) % n_host], host_tgts[(s + 1) % n_host])
                if pending is not None:
                    pending[1].synchronize()
                    loss_host = float(pending[0])
                pending = (buf, done)
                if s + 1 == n_steps:
                    done.synchronize()
                    loss_host = float(buf)
            else:
                loss = trainer.step(dev_imgs[i], dev_tgts[i], config=cfg)
            f, a = flops_per_image(cfg)
            flops += 3.0 * f * B
            attn_flops += 3.0 * a * B
            per_step.append((time.perf_counter() - t_host0) * 1e3)
        e1.record()
        # host enqueue time per step (no sync yet).  The mean includes launch-queue back-pressure once
        # the host runs ~1000 launches ahead; the minimum is the unblocked cost of enqueueing one step.
        host_ms[0] = {"mean": sum(per_step) / max(len(per_step), 1), "min": min(per_step) if per_step else 0.0}
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), flops, attn_flops, _lib.LAUNCHES[0] - launches0, loss_host

    # allocator priming (initialisation, not a step of the workload): the largest and the smallest
    # subnet once each, so the caching allocator owns its pools before any timed or warm-up step
    # (the search space changes every activation shape from step to step).
    prime = random.Random(12345)
    for e in sorted(ss["embed_dim"], reverse=True):
        for pick in (max, min):
            d = pick(ss["depth"])
            trainer.step(dev_imgs[0], dev_tgts[0], config={"layer_num": d, "embed_dim": [e] * d,
                                                           "num_heads": [pick(ss["num_heads"])] * d,
                                                           "mlp_ratio": [pick(ss["mlp_ratio"])] * d})
        trainer.step(dev_imgs[0], dev_tgts[0], rnd=prime)      # one mixed (per-layer h, r) subnet
    torch.cuda.synchronize()
    segments0 = torch.cuda.memory_stats().get("segment.all.allocated", 0)
    timed(cfg_warm, False)                                 # warm-up (untimed)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()                                    # samples clocks through BOTH timed regions
    ms, flops, attn_flops, launches, _ = timed(cfg_timed, False)
    host_enqueue_ms = host_ms[0]
    timed(cfg_warm[:3], True)
    ms_e2e, _, _, _, last_loss = timed(cfg_timed, True)
    clocks = sampler.stop() if sampler else None
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "config": args.config, "n_gpus": world, "value": B * world * K / (ms * 1e-3),
                              "ms_per_step": ms / K, "e2e_ms_per_step": ms_e2e / K, "gpu_launches": launches,
                              "host_enqueue_ms_per_step": host_enqueue_ms, "pdl": os.environ.get("CREAM_PDL", "1"),
                              "last_loss": last_loss, "clocks": clocks}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel roofline pass (instrumented; separate from the timed regions, same configs) ----
    # Every launch is bracketed by CUDA events on its own stream.  So that the events time the kernel
    # and not the host's enqueue gaps, each profiled step starts behind a spin kernel long enough for
    # the host to queue the step's launches before the GPU starts draining them.
    # The native runtime enqueues a whole forward / backward from C++, so the instrumented pass drives the
    # SAME kernels (same launch sequence, bit-identical results: tests/test_gpu_native.py) through the
    # Python sequencing, where every launch can be bracketed.
    prof_trainer = trainer if args.python_engine else SupernetTrainer(model, ss, native=False)
    ops.PROFILE = []
    spin = int(25e-3 * 1.9e9)
    for s, cfg in enumerate(cfg_timed[:4]):
        torch.cuda._sleep(spin)
        prof_trainer.forward_backward(dev_imgs[s % n_host], dev_tgts[s % n_host], config=cfg)
        torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg = {}
    for kind, a, b, fl, by in prof:
        d = agg.setdefault(kind, [0.0, 0.0, 0.0, 0])
        d[0] += a.elapsed_time(b) * 1e-3
        d[1] += fl
        d[2] += by
        d[3] += 1
    # cross-check from CUPTI (torch.profiler): per-kernel-name GPU time of the NATIVE path over the same steps
    kernel_table = None
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as tp:
            for s, cfg in enumerate(cfg_timed[:4]):
                trainer.step(dev_imgs[s % n_host], dev_tgts[s % n_host], config=cfg)
            torch.cuda.synchronize()
        rows = [(e.key, e.count, float(getattr(e, "device_time_total", getattr(e, "cuda_time_total", 0.0)))) for e in tp.key_averages()]
        rows = [r for r in rows if r[2] > 0]
        tot = sum(r[2] for r in rows) or 1.0
        kernel_table = [{"kernel": k[:72], "launches": c, "us": round(t, 1), "share": round(t / tot, 4)}
                        for k, c, t in sorted(rows, key=lambda r: -r[2])[:14]]
        kernel_table.append({"kernel": "TOTAL (4 steps)", "launches": sum(r[1] for r in rows), "us": round(tot, 1), "share": 1.0})
    except Exception as e:   # noqa: BLE001
        kernel_table = [{"unavailable": repr(e)[:200]}]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md)"
    hbm = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    tfile = ROOT / "profiles" / "ncu_traffic.json"
    if tfile.exists():   # dram__bytes_read+write per launch from the committed `ncu --set full` capture
        tj = json.loads(tfile.read_text())
        gl = [v for k, v in tj.items() if k.startswith("gemm_bf16_kernel")]
        if gl:
            traffic = sum(v["dram_bytes_per_launch"] * v["launches"] for v in gl) / sum(v["launches"] for v in gl)
    gsec, gfl, _, gcnt = agg.get("gemm", [1e-9, 0, 0, 0])
    asec, afl, aby, acnt = agg.get("attn_fwd", [1e-9, 0, 0, 0])
    bsec, bfl, bby, bcnt = agg.get("attn_bwd", [1e-9, 0, 0, 0])
    roofline = {"kernel": "gemm_bf16_kernel (all sliced linears: fwd, dgrad, wgrad)", "bound": "tensor",
                "achieved": gfl / gsec / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": gfl / gsec / 1e12 / peak_tf, "traffic": traffic,
                "traffic_note": "avg DRAM bytes per launch over the gemm launches in profiles/ncu_traffic.json", "launches": gcnt,
                "avg_launch_us": gsec / max(gcnt, 1) * 1e6, "peak_source": peak_src,
                "timing": "CUDA events around each launch on its stream, launches pre-queued behind a spin kernel"}
    attn = {"kernel": "attn_fwd_kernel (fused QK^T + RPE gather + softmax + PV)", "bound": "hbm",
            "achieved": aby / asec / 1e9, "peak": hbm, "unit": "GB/s", "frac": aby / asec / 1e9 / hbm,
            "tflops": afl / asec / 1e12, "tflops_frac_of_peak": afl / asec / 1e12 / peak_tf, "launches": acnt,
            "avg_launch_us": asec / max(acnt, 1) * 1e6}
    attn_bwd = {"kernel": "attention backward (all launches of cream_attn_bwd)", "bound": "hbm",
                "achieved": bby / bsec / 1e9, "peak": hbm, "unit": "GB/s", "frac": bby / bsec / 1e9 / hbm,
                "tflops": bfl / bsec / 1e12, "launches": bcnt, "avg_launch_us": bsec / max(bcnt, 1) * 1e6}
    imgs = B * world * K
    value = imgs / (ms * 1e-3)
    line = {
        "metric": conf["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": conf["workload"], "baseline_config": args.config, "global_batch": B * world,
                   "per_gpu_batch": B, "parallelism": f"dp{world}",
                   "runtime": "python sequencing" if args.python_engine else "native (cream_vit_fwd/bwd, cream_adamw_step)",
                   "grad_allreduce": ("one fp32 all-reduce per layer, overlapped" if args.overlap == 0 else
                                      f"{args.overlap} fp32 all-reduce(s) per step, all but the last under the backward") if world > 1 else "none",
                   "config_stream": "sample_configs with random.Random(0), identical on all ranks; the SAME K "
                                    "configurations are used for value, e2e and the roofline pass",
                   "l2": "inputs + activations per step (> 5 GB) far exceed the 126 MB L2; 4 rotating batches"},
        "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * 3 * 224 * 224 * 4 + B * 8,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K, "last_loss": last_loss},
        "gpu_launches": launches,
        "allocator_segments_grown_after_priming": torch.cuda.memory_stats().get("segment.all.allocated", 0) - segments0,
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "clocks": clocks,
        "roofline": roofline,
        "attention": attn,
        "attention_bwd": attn_bwd,
        "byte_movers": {k: {"GB/s": agg[k][2] / agg[k][0] / 1e9, "frac_of_hbm_peak": agg[k][2] / agg[k][0] / 1e9 / hbm,
                            "launches": agg[k][3], "avg_launch_us": agg[k][0] / agg[k][3] * 1e6}
                        for k in ("ln_fwd", "ln_bwd", "cast_scale", "bias_grad") if k in agg and agg[k][0] > 0},
        "gpu_time_share": {k: v[0] / max(sum(x[0] for x in agg.values()), 1e-12) for k, v in agg.items()},
        "kernel_table_cupti": kernel_table,
        "model_tflops": flops * world / (ms * 1e-3) / 1e12,
        "model_tflops_per_gpu": flops / (ms * 1e-3) / 1e12,
        "model_flops_frac_of_peak": flops / (ms * 1e-3) / 1e12 / peak_tf,
        "attn_core_share_of_flops": attn_flops / max(flops, 1.0),
    }
    if world == 1:
        line["gpu_reference"] = gpu_reference_supernet(spec, cfg_warm[:3], cfg_timed[:max(4, K // 2)], dev_imgs, dev_tgts, B)
        if line["gpu_reference"] and "value" in line["gpu_reference"]:
            line["speedup_vs_gpu_reference"] = line["value"] / line["gpu_reference"]["value"]
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"], _ = cpu_baseline(2, 1, config=args.config)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
