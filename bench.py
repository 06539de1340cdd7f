#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (contract: see DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--precision fp32|tc16]

A "step" is one stage-1 generator step of SDFace-GAN's SDF generator on synthetic latents and cameras (BASELINE.json configs[1]:
64^2 rays x 24 samples, hash grid on, per-GPU batch 32): forward with `return_sdf` + `return_eikonal`
(/root/reference/im2scene/training_utils.py:408-410), the reference's generator losses without the discriminator
(non-saturating term on the thumbnail, eikonal, minimal surface -- sdf_losses.py), backward through compositing, field and hash
grid, and the Adam update.  One process per GPU; units (images) are sharded over ranks, gradients are averaged with NCCL (DDP).

Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference path (oracle/) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R, S, STYLE = 64, 24, 256
SAMPLES_PER_IMAGE = R * R * S
# algorithmic work of the field per sample: MACs = 32*256 + 3*256^2 + 256 + 272*256 + 768
FIELD_FLOP_FWD = 2 * (32 * 256 + 3 * 256 * 256 + 256 + 272 * 256 + 768)


def g_losses(thumb, sdf, eik):
    """Generator-side losses of the reference's stage-1 step, minus the discriminator (training_utils.py:419-441)."""
    import torch.nn.functional as F
    gan = F.softplus(-thumb.mean(dim=(1, 2, 3))).mean()              # g_nonsaturating_loss on a stand-in critic
    eikonal = 0.1 * ((eik.norm(dim=-1) - 1) ** 2).mean()             # eikonal_lambda = 0.1
    min_surf = 0.05 * torch.exp(-100.0 * sdf.abs()).mean()           # min_surf_lambda = 0.05, beta = 100
    return gan + eikonal + min_surf


# ----------------------------------------------------------------------------------------------------------------------
# clocks

class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        # NVML in-process (a sample every 10 ms: a 5-step timed region of 70 ms still gets several); nvidia-smi as the fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [(0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap")]
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = reasons_fn(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b, _ in bits])
                time.sleep(0.01)
            return
        except Exception:
            pass
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.rows.append([x.strip() for x in o])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU baseline (oracle port)

def cpu_reference_step(batch, threads):
    """One fwd+bwd of the same workload through the CPU restatement (oracle/): torch-CPU field + C hash grid."""
    import oracle
    from oracle import field_oracle as fo
    torch.set_num_threads(threads)
    rs = np.random.RandomState(0)
    offsets, pls = oracle.grid_offsets(**fo.NGP_GRID)
    W = STYLE
    lim = np.sqrt(6 / W) / 25
    def U(shape, a):
        return torch.from_numpy(rs.uniform(-a, a, shape).astype(np.float32)).requires_grad_(True)
    p = {"network.encoder.embeddings": U((int(offsets[-1]), 2), 1e-4), "network.encoder.offsets": torch.from_numpy(offsets),
         "network.input_linear.weight": U((W, 32), np.sqrt(6 / 32) / 25), "network.input_linear.bias": U((W,), np.sqrt(1 / 32)),
         "network.sigma_linear.weight": U((1, W), lim), "network.sigma_linear.bias": U((1,), 1 / 16),
         "network.rgb_linear.weight": U((3, W), lim), "network.rgb_linear.bias": U((3,), 1 / 16),
         "sigmoid_beta": torch.tensor([0.1], requires_grad=True)}
    for name, k in [("network.pts_linears.%d" % i, W) for i in range(3)] + [("network.views_linears", W + 16)]:
        p[name + ".weight"] = U((W, k), 1 / 3 if name.endswith(".0") else np.sqrt(6 / k) / 25)
        p[name + ".bias"] = U((W,), np.sqrt(1 / k))
        for hb, kind in ((name + ".gamma", 1.0), (name + ".beta", 1.0)):
            p[hb + ".weight"] = U((W, STYLE), 0.05)
            p[hb + ".bias"] = U((W,), 1 / 16)
    import importlib
    sg = importlib.import_module("sdface-gan_b200.sdf_utils")
    cam, focal, near, far, _ = sg.generate_camera_params(R, "cpu", batch=batch)
    style = torch.randn(batch, STYLE)
    t_rand = torch.rand(batch, R, R)

    def step():
        for v in p.values():
            v.grad = None
        rgb, _, sdf, _, _, eik = fo.render(p, cam, focal, near, far, style, res=R, S=S, t_rand=t_rand, output_features=False,
                                           return_sdf=True, return_eikonal=True)
        loss = g_losses(rgb, sdf, eik)
        loss.backward()
        return float(loss.detach())
    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = args.ref_batch
    step = cpu_reference_step(batch, threads)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = batch / dt
    sample = "%d image(s) (%d samples) per step of the configs[1] workload, fwd+bwd, %d steps" % (batch, batch * SAMPLES_PER_IMAGE, args.steps)
    line = {"impl": "reference", "metric": "images/sec (generator fwd+bwd, field+composite)", "value": val, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "msamples_per_s": val * SAMPLES_PER_IMAGE / 1e6,
            "config": {"workload": "configs[1]: 64^2 SDF + hash-grid generator forward+backward (stage-1 G step), CPU sample", "rays": R,
                       "samples_per_ray": S, "batch_per_step": batch},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm

def run_ours(args):
    import torch.distributed as dist
    import sdface_gan_b200 as sg
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)
    B = args.batch
    mo, ro = sg.default_options("ngp", renderer_res=R, n_samples=S, perturb=1.0, no_features_output=True, return_sdf=True)
    g = sg.Generator(mo, ro, full_pipeline=False).to(dev)
    g.renderer.network.precision = args.precision
    if world > 1:      # identical weights on every rank
        for p_ in g.parameters():
            dist.broadcast(p_.data, 0)
    # The hash-table gradient (50.6 of the 54.6 MB exchanged) is all-reduced by the field's backward node itself, between the table
    # scatter and the weight-gradient kernels on the same stream (sdf_model._field.backward); SDFG_EARLY_EXCHANGE=0 leaves it to
    # DistributedDataParallel's last bucket instead (nothing left to overlap with: +0.7-1.0 ms per step in round 1).
    model = sg.distributed.data_parallel(g, device_ids=[local], early_table_exchange=os.environ.get("SDFG_EARLY_EXCHANGE", "1") == "1") if world > 1 else g
    opt = torch.optim.Adam(g.parameters(), lr=2e-5, betas=(0.0, 0.9), fused=True)      # im2scene/config.py:196-204 (stage 1); one fused update kernel

    # synthetic inputs: resident copies for `value`, pinned host copies for `e2e`
    cam, focal, near, far, _ = sg.generate_camera_params(R, dev, batch=B)
    z = torch.randn(B, STYLE, device=dev)
    host = [t.cpu().pin_memory() for t in (z, cam, focal, near, far)]
    h2d = sum(t.numel() * t.element_size() for t in host)

    def step(zz, cc, ff, nn_, fa):
        opt.zero_grad(set_to_none=True)
        _, thumb, sdf, eik = model([zz], cc, ff, nn_, fa, return_sdf=True, return_eikonal=True)
        loss = g_losses(thumb, sdf, eik)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step(z, cam, focal, near, far)
    barrier()

    # --- device-resident throughput + per-kernel roofline timing of the dominant kernel
    dominant = "gemm"                                   # every GEMM launch of the field (fwd, dgrad, wgrad)
    sg._lib.launch_count_reset()
    sg._lib.prof_enable(True, dominant)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            step(z, cam, focal, near, far)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = sg._lib.launch_count()
    sg._lib.prof_enable(False, "")
    k_ms, k_n = sg._lib.prof_collect()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # --- end to end: pinned host inputs -> device every step, loss read back every step
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dv = [h.to(dev, non_blocking=True) for h in host]
        loss = step(*dv)
        _ = loss.item()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())

    # --- inference forward (configs[2], field part: rays -> encoder -> field chain -> compositing), same per-GPU batch
    g_inf = sg.Generator(mo, ro, full_pipeline=False, ema=True).to(dev).eval()
    g_inf.load_state_dict(g.state_dict())
    g_inf.renderer.network.precision = args.precision
    mo_f, ro_f = sg.default_options("ngp", renderer_res=R, n_samples=S, perturb=0.)
    g_feat = sg.Generator(mo_f, ro_f, full_pipeline=False, ema=True).to(dev).eval()        # with the 256-channel feature map
    g_feat.renderer.network.precision = args.precision
    inf = {}
    with torch.no_grad():
        for name, gi in (("thumb_only", g_inf), ("with_features", g_feat)):
            for _ in range(3):
                gi([z], cam, focal, near, far)
            barrier()
            sg._lib.prof_enable(True, dominant)
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            i0.record()
            for _ in range(args.steps):
                gi([z], cam, focal, near, far)
            i1.record()
            barrier()
            sg._lib.prof_enable(False, "")
            ik_ms, ik_n = sg._lib.prof_collect()
            t = torch.tensor([i0.elapsed_time(i1) / args.steps], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ims = float(t.item())
            inf[name] = {"ms_per_pass": ims, "images_per_s": world * B / (ims * 1e-3), "msamples_per_s": world * B * SAMPLES_PER_IMAGE / (ims * 1e-3) / 1e6,
                         "field_chain_ms": ik_ms / max(ik_n, 1),
                         "field_chain_tflops": B * SAMPLES_PER_IMAGE * FIELD_FLOP_FWD / (ik_ms / max(ik_n, 1) * 1e-3) / 1e12 if ik_n else None}

    # --- configs[2]: the 256^2 generator forward (renderer 64^2 x 24 -> StyleGAN2 decoder -> 256^2), eval, B = 64 split over the GPUs
    #     (strong scaling: 64 / world images per GPU and pass); end to end from resident latents / cameras to the image tensor
    inf256 = None
    if args.precision == "tc16" and 64 % world == 0:
        try:
            Bi = 64 // world
            mo_d, ro_d = sg.default_options("ngp", size=256, renderer_res=R, n_samples=S, perturb=0.)
            g_full = sg.Generator(mo_d, ro_d, full_pipeline=True, ema=True).to(dev).eval()
            g_full.renderer.network.precision = args.precision
            cam_i, focal_i, near_i, far_i, _ = sg.generate_camera_params(R, dev, batch=Bi)
            z_i = torch.randn(Bi, STYLE, device=dev)
            with torch.no_grad():
                for _ in range(3):
                    g_full([z_i], cam_i, focal_i, near_i, far_i)
                barrier()
                sg._lib.prof_enable(True, "gemm")
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for _ in range(args.steps):
                    img, _thumb = g_full([z_i], cam_i, focal_i, near_i, far_i)
                f1.record()
                barrier()
                sg._lib.prof_enable(False, "")
                gk_ms, gk_n = sg._lib.prof_collect()
                graph_ms, lat, graph_err = None, {}, None
                try:
                    # the same pass replayed as a CUDA graph (sg.GraphedGenerator: the serving call for a fixed batch shape)
                    gg = sg.GraphedGenerator(g_full, [z_i], cam_i, focal_i, near_i, far_i)
                    for _ in range(2):
                        gg([z_i], cam_i, focal_i, near_i, far_i)
                    torch.cuda.synchronize()
                    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    q0.record()
                    for _ in range(args.steps):
                        gg([z_i], cam_i, focal_i, near_i, far_i)
                    q1.record()
                    torch.cuda.synchronize()
                    del gg
                    # single-image latency (the interactive / demo case): eager (launch-bound: ~230 launches) and as a graph
                    if rank == 0:
                        one = (cam_i[:1].contiguous(), focal_i[:1].contiguous(), near_i[:1].contiguous(), far_i[:1].contiguous())
                        g1 = sg.GraphedGenerator(g_full, [z_i[:1]], *one)
                        for name, fn in (("eager", lambda: g_full([z_i[:1]], *one)), ("graphed", lambda: g1([z_i[:1]], *one))):
                            for _ in range(3):
                                fn()
                            torch.cuda.synchronize()
                            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            l0.record()
                            for _ in range(20):
                                fn()
                            l1.record()
                            torch.cuda.synchronize()
                            lat[name] = l0.elapsed_time(l1) / 20
                        del g1
                    torch.cuda.synchronize()
                except Exception as e:            # the graph legs are extras: the eager numbers above must survive a capture failure
                    graph_err = "%s: %s" % (type(e).__name__, e)
                # renderer alone (same batch), to split the pass
                style_i = g_full.style(z_i)
                for _ in range(2):
                    g_full.renderer(cam_i, focal_i, near_i, far_i, styles=style_i)
                barrier()
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r0.record()
                for _ in range(args.steps):
                    g_full.renderer(cam_i, focal_i, near_i, far_i, styles=style_i)
                r1.record()
                barrier()
            t = torch.tensor([f0.elapsed_time(f1) / args.steps, r0.elapsed_time(r1) / args.steps, (q0.elapsed_time(q1) / args.steps) if graph_err is None else 0.0],
                             device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            full_ms, rend_ms, graph_ms = float(t[0]), float(t[1]), (float(t[2]) if float(t[2]) > 0 else None)
            # decoder MACs per image: 3x3 modulated convolutions 256->512@64^2, 512->256 (x2 up), 256->256@128^2, 256->128 (x2 up), 128->128@256^2
            dec_flop = 2 * 9 * (64 * 64 * 256 * 512 + 64 * 64 * 512 * 256 + 128 * 128 * 256 * 256 + 128 * 128 * 256 * 128 + 256 * 256 * 128 * 128)
            inf256 = {"workload": "configs[2]: ffhq_256_sdf_ngp generator forward (renderer + decoder), eval, B = 64 over %d GPU(s)" % world,
                      "batch_per_gpu": Bi, "ms_per_pass": full_ms, "images_per_s": 64 / (full_ms * 1e-3), "renderer_ms": rend_ms,
                      "decoder_ms": full_ms - rend_ms, "scaling": "strong",
                      "graphed_ms_per_pass": graph_ms, "graphed_images_per_s": (64 / (graph_ms * 1e-3)) if graph_ms else None, "graph_error": graph_err,
                      "latency_ms_batch1": lat,
                      "gemm_kernel_ms_per_pass": gk_ms / args.steps, "gemm_kernel_launches_per_pass": gk_n / args.steps,
                      "generator_tflops_algorithmic": Bi * (SAMPLES_PER_IMAGE * FIELD_FLOP_FWD + dec_flop) / (full_ms * 1e-3) / 1e12,
                      "image_shape": list(img.shape)}
        except Exception as e:        # configs[2] is an extra object of the line: the headline numbers above must survive it
            inf256 = {"error": "%s: %s" % (type(e).__name__, e)}
            g_full = None
        del g_full
        torch.cuda.empty_cache()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md sustained 1.4 PFLOP/s)"
        N = B * SAMPLES_PER_IMAGE
        # GEMM flops per step: forward + eikonal dgrad of the trunk + backward (dgrad + wgrad), per sample
        trunk = 2 * (3 * 256 * 256)
        first = 2 * 32 * 256
        views = 2 * 272 * 256
        flop_step = N * ((first + trunk + views)                # forward
                         + (trunk + first)                      # eikonal pass: trunk dgrads + input-linear dgrad
                         + (2 * trunk + 2 * first + views + 2 * 256 * 256 + 2 * 16 * 256))   # backward: dgrad + wgrad
        achieved = flop_step / (k_ms / args.steps * 1e-3) / 1e12 if k_n else None
        # what the kernels EXECUTE is less: input_linear is folded into the first FiLM layer (its x part as hi + lo pairs), every layer
        # carries one K = 16 step for the folded FiLM offset.  MMA K per layer: 112 | 272 | 272 | 288 (forward); 2 trunk dgrads + the
        # N = 32 input stage (eikonal); 3 dgrads + input stage + weight gradients with K = 288 | 272 | 272 | 48 (backward).
        flop_exec = N * 2 * 256 * ((112 + 272 + 272 + 288) + (2 * 256 + 32) + (3 * 256 + 32) + (288 + 272 + 272 + 48))
        traffic, traffic_note = None, None
        try:        # dram bytes of the same launches per step, from the committed ncu --set full capture (B = 32)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            if B == tj.get("batch") and args.precision == "tc16":
                traffic, traffic_note = tj["dram_bytes_per_step"], tj["note"]
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "field GEMMs (%s)" % ("gemm_f32_kernel, fp32 SIMT" if args.precision == "fp32" else "tcgen05"),
                "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": (achieved / tensor_peak) if achieved else None,
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "kernel_ms_per_step": k_ms / args.steps, "kernel_launches_per_step": k_n / args.steps,
                "kernel_share_of_step": (k_ms / args.steps) / ms,
                "algorithmic_flop_per_step": flop_step, "executed_flop_per_step": flop_exec if args.precision == "tc16" else flop_step,
                "achieved_executed": (flop_exec / (k_ms / args.steps * 1e-3) / 1e12) if (k_n and args.precision == "tc16") else achieved,
                # the same launches against the OTHER roof: measured dram bytes (traffic) / their summed time / the measured HBM peak.  The
                # step saves every layer's activations for the backward, so its GEMM kernels sit closer to the HBM roof than to the tensor one.
                "hbm_GBps": (traffic / (k_ms / args.steps * 1e-3) / 1e9) if (traffic and k_n) else None,
                "hbm_peak_GBps": peaks.get("hbm_gbs"),
                "hbm_frac": (traffic / (k_ms / args.steps * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6547.0)) if (traffic and k_n) else None,
                "note": "achieved = algorithmic GEMM flop of the reference network for this step (SURVEY 8d: forward 550 912 flop/sample incl. heads' "
                        "GEMM part, eikonal dgrads, backward dgrad + wgrad) / summed CUDA-event time of the tcgen05 launches of the step"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cstep = cpu_reference_step(1, threads)
            cstep()
            t0 = time.perf_counter()
            n = 0
            while n < 3 or (time.perf_counter() - t0 < 10 and n < 50):
                cstep()
                n += 1
            cdt = (time.perf_counter() - t0) / n
            cpu = {"value": 1.0 / cdt, "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": "1 image (98304 samples) per step, fwd+bwd, %d steps" % n}
        line = {"metric": "images/sec (generator fwd+bwd, field+composite)", "value": world * B / (ms * 1e-3), "unit": "images/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f16 operands (activations, loss-scaled gradients), f32 accumulate", "data": "synthetic",
                "msamples_per_s": world * N / (ms * 1e-3) / 1e6,
                "config": {"workload": "configs[1]: 64^2 SDF + hash-grid (ngp=1) generator forward+backward, stage-1 G step", "rays": R,
                           "samples_per_ray": S, "batch_per_gpu": B, "global_batch": world * B, "parallelism": "dp%d" % world,
                           "l2": "per-step saved activations + gradient tiles (%.1f GB) exceed the 126 MB L2; no flush needed" % (N * 5.3e3 / 1e9)},
                "clocks": clk.summary(),
                "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "roofline": roof, "inference": inf, "inference_256": inf256, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=1, help="images per CPU reference step (bounded sample)")
    ap.add_argument("--precision", default=os.environ.get("SDFG_PRECISION", "tc16"), choices=["fp32", "tc16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
