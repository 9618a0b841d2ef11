#!/usr/bin/env python
"""bench.py -- CLIP-prefix LM step on B200 (BASELINE.json metric: mapper train samples/s; few-shot answers/s).

    python bench.py --gpus 1 --steps 20 --warmup 3                      # this repo's CUDA path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W                       # data-parallel, weak scaling (256 samples / GPU)
    python bench.py --impl reference --steps K --warmup W               # the reference's own CPU path on the host cores

One JSON line on stdout (rank 0).  A "step" = forward + backward-to-the-mapper + caption cross-entropy of one
synthetic Conceptual-Captions batch (BASELINE.json configs[1]: GPT-2 small, transformer mapper, batch 256 per
GPU, bf16 tensor-core math with fp32 accumulate), the NCCL all-reduce of the mapper gradients when N > 1, and the
fused AdamW update of the mapper.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# ---- workload ------------------------------------------------------------------------------------
C2 = dict(model_version="gpt2", mapping_type="transformer", prefix_length=10, clip_length=10, clip_dim=512, num_layers=8,
          batch_per_gpu=256, text_len=40, vocab=50257)
C4 = dict(model_version="gpt2-medium", mapping_type="mlp", prefix_length=10, clip_length=10, clip_dim=512, num_layers=8,
          batch=128, num_shots=4, max_length=10)
C5 = dict(model_version="gpt2-xl", mapping_type="mlp", prefix_length=10, clip_length=10, clip_dim=768, num_layers=8,
          batch_per_gpu=64, text_len=40, vocab=50257)
GFLOP_PER_SAMPLE_C2 = 27.88          # algorithmic work per sample: LM 23.30 + transformer mapper 4.58
GFLOP_PER_SAMPLE = {"c2": 27.88, "c5": 309.7}       # c5: LM 308.9 + MLP mapper 0.79
WORKLOAD_TEXT = {
    "c2": "Conceptual Captions mapper training step, BASELINE configs[1]: frozen GPT-2 small (d=768, L=12, V=50257), "
          "transformer mapper (8 layers, 8 heads, clip_length 10), prefix 10, text 40 (T=50), synthetic CLIP ViT-B/32 512-d embeddings",
    "c5": "Large-LM scaling, BASELINE configs[4]: frozen GPT-2 XL (d=1600, L=48, H=25, V=50257), MLP mapper, prefix 10, text 40 "
          "(T=50), synthetic CLIP ViT-L/14 768-d embeddings, 64 samples per GPU (512 on 8 GPUs)",
}
WORKLOAD = "c2"
ALLREDUCE_NOTE = "none (one GPU)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops=float(d["bf16_tflops_sustained"]), tflops_burst=float(d["bf16_tflops"]), hbm=float(d["hbm_gbs"]),
                    source="MEASURED_PEAKS.json (sustained bf16 GEMM, kernel timed inside a long step)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback of B200_PROFILING.md (file absent)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---- the reference's CPU implementation of the path, used for --impl reference and cpu_baseline ----------------------
C1 = dict(model_version="gpt2", mapping_type="mlp", prefix_length=10, clip_length=10, clip_dim=512, num_layers=8,
          batch=8, text_len=40, vocab=50257)          # BASELINE configs[0]: the reference's own CPU-runnable case
REF_ARM_BATCH = 64                                     # --impl reference: bounded sample of the 256-caption step


def cpu_reference(W, batch, steps: int, warmup: int):
    """Time forward + backward of the caption step on the host cores, in a child process that sees NO GPU
    (CUDA_VISIBLE_DEVICES=""): the reference module picks its device from torch.cuda.is_available() at import
    (clipcap.py:23), and it prints its architecture to stdout, which must not reach this process' one JSON line."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    spec = json.dumps({"W": W, "batch": batch, "steps": steps, "warmup": warmup})
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-reference-worker", spec], env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("CPU reference worker failed:\n" + r.stderr[-2000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def cpu_reference_worker(spec: str):
    """Runs the UNMODIFIED reference module (``ClipCaptionPrefix`` of clipcap.py, imported from oracle/_ref -- see
    oracle/install_reference.py) when its file travelled with the snapshot (kind "reference"), else the pinned oracle
    port (kind "port").  Median over `steps`."""
    a = json.loads(spec)
    W, batch, steps, warmup = a["W"], a["batch"], a["steps"], a["warmup"]
    real_stdout = os.dup(1)
    os.dup2(2, 1)                      # everything the reference prints goes to stderr
    from oracle import clip_prefix_lm as orc
    from oracle import reference_shim
    import eavqa_b200.synthetic as syn
    torch.set_num_threads(os.cpu_count() or 1)
    cfg_lm = syn.lm_config(W["model_version"], vocab=W["vocab"])
    lm_w = syn.make_lm_weights(cfg_lm, seed=0)
    mapper_w = syn.make_mapper_params(W["mapping_type"], W["clip_dim"], cfg_lm["d_model"], W["prefix_length"], W["clip_length"],
                                      W["num_layers"], seed=1, perturb_norm=True)
    b = syn.make_caption_batch(batch, W["text_len"], W["clip_dim"], W["vocab"], seed=2021)
    kind = "reference" if reference_shim.reference_dir() is not None else "port"
    if kind == "reference":
        clipcap, _, GPT2Config, holder = reference_shim.import_reference()
        ref = reference_shim.build_reference_model(clipcap, GPT2Config, holder, cfg_lm, lm_w, mapper_w,
                                                   prefix_length=W["prefix_length"], clip_length=W["clip_length"],
                                                   clip_dim=W["clip_dim"], num_layers=W["num_layers"], mapping_type=W["mapping_type"])

        def one():
            ref.zero_grad(set_to_none=True)
            out = ref(question_tokens=b["input_ids"], labels=b["labels"], prefix=b["clip_embeddings"],
                      question_mask=b["attention_mask"], pad_token_id=50256)
            out.loss.backward()
            return float(out.loss)
        what = "the unmodified reference module (clipcap.ClipCaptionPrefix over HF GPT2LMHeadModel)"
    else:
        cfg = dict(n_layer=cfg_lm["n_layer"], n_head=cfg_lm["n_head"], d_model=cfg_lm["d_model"], prefix_length=W["prefix_length"],
                   clip_length=W["clip_length"], mapping_type=W["mapping_type"], num_layers=W["num_layers"])

        def one():
            return orc.train_step(lm_w, mapper_w, cfg, b["input_ids"], b["clip_embeddings"], b["attention_mask"], b["labels"])[0]
        what = "the pinned oracle port (oracle/clip_prefix_lm.py; the reference's file did not travel to this box)"
    times, loss = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = one()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    res = dict(value=batch / sec, unit="samples/s", cores=torch.get_num_threads(), kind=kind, ms_per_step=sec * 1e3, loss=loss,
               sample="%s, fp32, %s mapper, GPT-2 small, batch %d, text 40: zero_grad + forward + loss.backward(), %d warm-up + "
                      "%d timed steps, median" % (what, W["mapping_type"], batch, warmup, steps))
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(res), flush=True)


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation on THIS arm's configuration (configs[1]: transformer mapper,
    GPT-2 small), each step a bounded sample of the 256-caption batch."""
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    r = cpu_reference(C2, REF_ARM_BATCH, steps, warmup)
    cfg = workload_config(1, note="reference CPU path on the host cores; every step is a bounded sample of the workload: "
                                  "%d of the 256 captions of a step (a CPU step of 256 takes ~4x as long at the same rate)" % REF_ARM_BATCH)
    cfg.update(batch_per_gpu=REF_ARM_BATCH, global_batch=REF_ARM_BATCH, parallelism="cpu", optimizer="none (forward + backward only)")
    line = {"impl": "reference", "metric": "mapper_train_samples_per_sec", "value": r["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "loss": r["loss"], "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(world, note=""):
    W = C5 if WORKLOAD == "c5" else C2
    return {"workload": WORKLOAD_TEXT[WORKLOAD], "batch_per_gpu": W["batch_per_gpu"], "global_batch": W["batch_per_gpu"] * world,
            "allreduce": ALLREDUCE_NOTE,
            "parallelism": "dp%d" % world, "optimizer": "fused AdamW on the mapper (in the timed region)",
            "l2": "inputs larger than L2: ~3 GB of activations per step >> 126 MB L2, no flush needed", "note": note}


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--cpu-reference-worker":
        cpu_reference_worker(sys.argv[2])
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-generate", action="store_true", help="skip the few-shot answers/s leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rices", action="store_true", help="skip the RICES retrieval leg (SURVEY.md 8f row 4)")
    ap.add_argument("--profile-report", default="", help="write the per-shape GEMM timing report to this file")
    ap.add_argument("--only-timed", action="store_true", help="run only warm-up + the timed region (for ncu launch lists)")
    ap.add_argument("--overlap-allreduce", action="store_true",
                    help="N > 1: bucketed gradient all-reduce overlapped with the mapper backward instead of one all-reduce after the step")
    ap.add_argument("--exchange", default="fused", choices=["fused", "fused-overlap", "nccl"],
                    help="N > 1: 'fused' = ONE kernel per rank does reduce-scatter + AdamW + all-gather over NVLink peer / multicast "
                         "memory (eavqa_sharded_adamw_step); 'fused-overlap' = the same kernel per gradient bucket on a communication "
                         "stream while the mapper backward still runs; 'nccl' = NCCL all-reduce of the flat gradient, then the fused AdamW on every rank")
    ap.add_argument("--nccl-max-ctas", type=int, default=0, help="N > 1: cap on NCCL's CTAs per collective (0 = NCCL's default)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the headline): 256 samples per GPU; strong: 256 samples in total, 256 / N per GPU (SURVEY.md 8d)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2 = BASELINE configs[1] (the metric's configuration, default); c5 = configs[4], GPT-2 XL, 64 samples / GPU")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    W = C5 if WORKLOAD == "c5" else C2
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU path (use --impl reference for the CPU port)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        # NCCL prints its version line to STDOUT when the first communicator is created: keep stdout for the JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            pg_options = None
            if args.nccl_max_ctas > 0:
                # cap the SMs NCCL's kernels may take, so that a collective overlapped with the step's own kernels does not
                # starve them (round 1: with the default channel count the bucketed overlap LOST at N = 8)
                pg_options = dist.ProcessGroupNCCL.Options()
                pg_options.config.max_ctas = args.nccl_max_ctas
                pg_options.config.min_ctas = min(args.nccl_max_ctas, 4)
            dist.init_process_group("nccl", device_id=dev, pg_options=pg_options)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    import eavqa_b200
    import eavqa_b200.synthetic as syn
    from eavqa_b200 import lib
    from eavqa_b200.optim import FlatAdamW
    L = lib.load()

    lm_cfg = syn.lm_config(W["model_version"], vocab=W["vocab"])
    lm_w = syn.make_lm_weights(lm_cfg, seed=0)
    model = eavqa_b200.ClipCaptionPrefixB200(prefix_length=W["prefix_length"], clip_length=W["clip_length"],
                                             prefix_size=W["clip_dim"], num_layers=W["num_layers"],
                                             mapping_type=W["mapping_type"], model_version=W["model_version"],
                                             lm_state_dict=lm_w)
    # mapper weights: the seeded generator of the parity cases (oracle/cases.py), so that rank 0's first batch of the c2
    # workload IS the case `train_c2_full_b256` whose reference loss is committed under tests/golden/
    model.clip_project.load_state_dict(syn.make_mapper_params(W["mapping_type"], W["clip_dim"], lm_cfg["d_model"], W["prefix_length"],
                                                              W["clip_length"], W["num_layers"], seed=1, perturb_norm=True))
    model = model.to(dev).train()
    B = W["batch_per_gpu"]
    if args.scaling == "strong":
        assert B % world == 0, "strong scaling needs the global batch to divide by the number of GPUs"
        B = B // world
    host = syn.make_caption_batch(B, W["text_len"], W["clip_dim"], W["vocab"], seed=2021 + rank)
    host = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    from eavqa_b200.parallel import NvlinkShardedAdamW, OverlappedGradReducer
    global ALLREDUCE_NOTE
    opt, fused = None, False
    if world > 1 and args.exchange.startswith("fused") and not args.overlap_allreduce:
        # the exchange step and the optimiser as ONE kernel per rank over NVLink / NVSwitch peer memory (csrc/collective.cu).
        # Measured (profiles/README.md, round 2): N = 2 0.277 ms against 0.535 ms for NCCL all-reduce + AdamW.
        try:
            opt = NvlinkShardedAdamW(model, lr=1e-4, overlap=args.exchange == "fused-overlap")
            fused = True
            ALLREDUCE_NOTE = ("none: reduce-scatter + AdamW + all-gather fused in one kernel per rank (%s), barriers inside the kernel%s"
                              % ("NVLS multimem.ld_reduce / multimem.st" if opt.multicast else "NVLink peer loads / stores",
                                 "; per gradient bucket on a communication stream during the mapper backward" if opt.overlap else ""))
        except Exception as e:      # symmetric memory unavailable on this box: the NCCL path below, and the line says so
            print("[bench] fused exchange unavailable (%s: %s); using the NCCL all-reduce" % (type(e).__name__, e), file=sys.stderr)
            opt = None
    if opt is None:
        opt = FlatAdamW(model, lr=1e-4)
    # --overlap-allreduce: bucketed all-reduce on a communication stream, started per pair of mapper layers while the rest of
    # the mapper backward runs.  Measured (profiles/README.md): N = 2 11.20 vs 11.34 ms/step for one all-reduce after the
    # step, N = 8 11.50 vs 11.26 ms/step (NVLS makes the 167 MB all-reduce cost only ~0.4 ms; NCCL's CTAs slow the GEMMs they
    # overlap by about as much).
    reducer = OverlappedGradReducer(model) if (world > 1 and args.overlap_allreduce) else None
    if world > 1 and not fused:
        ALLREDUCE_NOTE = ("bucketed NCCL all-reduce overlapped with the mapper backward" if reducer is not None else
                          "one NCCL all-reduce of the flat fp32 mapper gradient after the step") + \
                         (", NCCL capped at %d CTAs" % args.nccl_max_ctas if args.nccl_max_ctas > 0 else "")

    def step(b):
        out = model(question_tokens=b["input_ids"], labels=b["labels"], prefix=b["clip_embeddings"], question_mask=b["attention_mask"])
        out.loss.backward()
        if fused:
            opt.step()                                       # sum over ranks, 1/W, AdamW, broadcast: one kernel
        else:
            g = model.last_flat_grads
            if reducer is not None:
                reducer.reduce(g)
            elif world > 1:
                dist.all_reduce(g)                           # sum; the 1/W of the DDP mean is folded into AdamW
            opt.step(g, grad_scale=1.0 / world)
        opt.zero_grad()
        return out.loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms) / n

    # ---- parity guard: the loss of the step's first forward against the UNMODIFIED reference's loss on the same inputs
    #      (tests/golden/train_c2_full_b256.json, written by oracle/validate_against_reference.py); bar 1e-3 relative
    loss_check = None
    fixture = os.path.join(ROOT, "tests", "golden", "train_c2_full_b256.json")
    with torch.no_grad():
        loss0 = float(model(question_tokens=resident["input_ids"], labels=resident["labels"], prefix=resident["clip_embeddings"],
                            question_mask=resident["attention_mask"]).loss)
    assert loss0 == loss0 and abs(loss0) < 1e4, "bench.py: the step's loss is not finite (%r)" % loss0
    if WORKLOAD == "c2" and B == C2["batch_per_gpu"] and rank == 0 and os.path.exists(fixture):
        with open(fixture) as f:
            ref_loss = json.load(f)["loss"]
        rel = abs(loss0 - ref_loss) / abs(ref_loss)
        loss_check = {"loss": loss0, "reference_loss": ref_loss, "rel_err": rel, "bar": 1e-3,
                      "fixture": "tests/golden/train_c2_full_b256.json (the reference module's fp32 loss on these 256 captions)"}
        assert rel <= 1e-3, "bench.py: loss %.6f differs from the reference's %.6f by %.2e (> 1e-3)" % (loss0, ref_loss, rel)
    else:
        loss_check = {"loss": loss0, "reference_loss": None, "note": "no committed reference loss for this workload / shard"}

    # ---- kernel-resident throughput: inputs already in HBM ---------------------------------------------------------
    for _ in range(args.warmup):
        step(resident)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.eavqa_launch_count()
    ms_step = timed(lambda: step(resident), args.steps)
    launches = (L.eavqa_launch_count() - launches0) // args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = B * world / (ms_step * 1e-3)
    if args.only_timed:
        if rank == 0:
            print(json.dumps({"metric": "mapper_train_samples_per_sec", "value": value, "ms_per_step": ms_step,
                              "gpu_launches": int(launches), "note": "--only-timed (profiling run; not a bench value)"}), flush=True)
        return

    # ---- end to end through the public API: pinned host batch -> H2D -> step -> loss read back ------------------------
    def e2e_step():
        b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        loss = step(b)
        return float(loss.detach())                          # D2H of the step's result
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    e2e = {"value": B * world / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
           "ms_per_step": ms_e2e}

    # ---- N > 1: the exchange step + optimiser on its own (gradients already final), against the NVLink bytes it moves ------
    exchange = None
    if fused:
        for _ in range(3):
            opt.step(opt.grads)
        ms_x = timed(lambda: opt.step(opt.grads), 10)
        assert not opt.timed_out(), "bench.py: a barrier inside the fused exchange kernel timed out (a rank fell out of step)"
        gb = opt.link_bytes_per_direction() / 1e9
        exchange = {"kernel": "sharded_adamw_kernel (reduce-scatter + AdamW + all-gather, one launch per rank)", "ms": ms_x,
                    "path": "NVLS multimem.ld_reduce / multimem.st" if opt.multicast else "NVLink peer loads / stores",
                    "link_gbytes_per_direction": gb, "achieved_gb_s_per_direction": gb / (ms_x * 1e-3),
                    "nominal_peak_gb_s_per_direction": 900.0,
                    "note": "timed alone with CUDA events, max over ranks; bytes per GPU and direction (parallel.NvlinkShardedAdamW.link_bytes_per_direction)"}

    # ---- roofline of the dominant kernel (tcgen05 GEMM): per-launch CUDA events, outside the timed region --------------
    import ctypes as C
    pk = peaks()
    prof_steps = 3
    lib.check(L.eavqa_profile_begin())
    for _ in range(prof_steps):
        step(resident)
    tot_ms, tot_fl, nl = C.c_double(), C.c_double(), C.c_int64()
    rep = C.create_string_buffer(1 << 16)
    lib.check(L.eavqa_profile_end(C.byref(tot_ms), C.byref(tot_fl), C.byref(nl), rep, len(rep)))
    gemm_ms, gemm_tf = tot_ms.value / prof_steps, tot_fl.value / prof_steps / 1e12
    achieved = gemm_tf / (gemm_ms * 1e-3)
    step_tflops = GFLOP_PER_SAMPLE[WORKLOAD] * B / 1e3 / (ms_step * 1e-3)
    # DRAM traffic of the GEMM launches: ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the 196 GEMM launches
    # of one step, per launch like `achieved` (profiles/r02_gemm_traffic_v1.json, made by tools/ncu_step_summary.py from
    # the committed launch list); algorithmic bytes per launch from the same per-launch records as the timings.
    import re
    alg = [float(m.group(2)) * int(m.group(1)) for m in re.finditer(r"launches=(\d+) .*alg_mbytes=([0-9.]+)", rep.value.decode())]
    alg_gb_per_launch = sum(alg) / 1e3 / max(nl.value, 1)
    traffic, traffic_note = None, "no ncu traffic summary under profiles/"
    tp = os.path.join(ROOT, "profiles", "r02_gemm_traffic_v1.json")
    if os.path.exists(tp) and WORKLOAD == "c2":
        with open(tp) as f:
            t = json.load(f)
        traffic = t["gemm_dram_bytes_per_launch"] / 1e9
        traffic_note = ("GB of DRAM read+write per GEMM launch, averaged over the %d GEMM launches of one step (ncu, %s); "
                        "algorithmic %.4f GB per launch (operands + outputs once; part of every output is still in the "
                        "126 MB L2 when its kernel ends)" % (t["gemm_launches_per_step"], os.path.basename(tp), alg_gb_per_launch))
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_tn_kernel (tcgen05/TMEM, all GEMMs of the step)", "achieved": achieved,
                "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": pk["source"], "gemm_ms_per_step": gemm_ms, "gemm_tflop_per_step": gemm_tf,
                "gemm_launches_per_step": nl.value // prof_steps, "gemm_share_of_step": gemm_ms / ms_step,
                "whole_step_tflops": step_tflops, "whole_step_frac": step_tflops / pk["tflops"],
                "algorithmic_gflop_per_sample": GFLOP_PER_SAMPLE[WORKLOAD]}
    if args.profile_report and rank == 0:
        with open(args.profile_report, "w") as f:
            f.write(rep.value.decode())

    line = {"metric": "mapper_train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": dict(workload_config(world), batch_per_gpu=B, global_batch=B * world),
            "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "loss_check": loss_check}
    if exchange is not None:
        line["exchange"] = exchange

    # ---- few-shot VQA answers/s (BASELINE configs[3]) and the CPU baseline: rank 0, N = 1 only ------------------------
    if reducer is not None:
        reducer.close()
    del model, opt, reducer
    torch.cuda.empty_cache()
    if WORKLOAD != "c2":
        args.no_generate = args.no_cpu_baseline = args.no_rices = True         # those legs belong to the metric's own configuration
    if world == 1 and not args.no_generate:
        line["few_shot_generate"] = bench_generate(dev, eavqa_b200, syn, args)
    if world == 1 and not args.no_rices:
        line["rices_retrieval"] = bench_rices(with_cpu=not args.no_cpu_baseline)
    if world == 1 and not args.no_cpu_baseline:
        # BASELINE configs[0] exactly: MLP mapper, batch 8, 3 warm-up + 10 timed steps, median
        r = cpu_reference(C1, C1["batch"], steps=10, warmup=3)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_generate(dev, eavqa_b200, syn, args):
    c = C4
    k = c["num_shots"]
    lm_cfg = syn.lm_config(c["model_version"], vocab=50257 + k + 1)
    lm_w = syn.make_lm_weights(lm_cfg, seed=0, hot_rows=512)
    torch.manual_seed(1)
    model = eavqa_b200.ClipCaptionPrefixB200(prefix_length=c["prefix_length"], clip_length=c["clip_length"], prefix_size=c["clip_dim"],
                                             num_layers=c["num_layers"], mapping_type=c["mapping_type"],
                                             model_version=c["model_version"], lm_state_dict=lm_w,
                                             special_token_id=50257 + k).to(dev).eval()
    host = syn.make_fewshot_batch(c["batch"], k, c["clip_dim"], lm_cfg["vocab"], 50257 + k, seed=2021, pad_token_id=50256)
    host = {kk: v.pin_memory() for kk, v in host.items()}

    def gen():
        b = {kk: v.to(dev, non_blocking=True) for kk, v in host.items()}
        # eos_token_id=None: every row decodes all max_length tokens (worst case; no early exit)
        return model.generate(question_tokens=b["input_ids"], prefix=b["clip_embeddings"], question_mask=b["attention_mask"],
                              max_length=c["max_length"], pad_token_id=50256, eos_token_id=None)
    model.gpt.config.eos_token_id = None
    for _ in range(3):
        gen()
    n = max(3, min(args.steps, 10))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = gen()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    T0 = host["input_ids"].shape[1] + 9 * (k + 1)

    # prefill / decode split: the same call with max_length = 1 is the prefill (+ the first pick); the rest is the nine
    # single-token steps.  Prefill is tensor-bound, a decode step HBM-bound: weights once + the KV history.
    def gen1():
        b = {kk: v.to(dev, non_blocking=True) for kk, v in host.items()}
        return model.generate(question_tokens=b["input_ids"], prefix=b["clip_embeddings"], question_mask=b["attention_mask"],
                              max_length=1, pad_token_id=50256, eos_token_id=None)
    for _ in range(2):
        gen1()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        gen1()
    e1.record()
    torch.cuda.synchronize()
    ms_prefill = e0.elapsed_time(e1) / n
    pk = peaks()
    L_, d_, B_ = lm_cfg["n_layer"], lm_cfg["d_model"], c["batch"]
    steps = c["max_length"] - 1
    ms_step = (ms - ms_prefill) / steps
    weight_bytes = L_ * 12 * d_ * d_ * 2 + ((lm_cfg["vocab"] + 63) // 64 * 64) * d_ * 2
    kv_bytes = sum(B_ * (T0 + i) * L_ * 2 * d_ * 2 for i in range(1, steps + 1)) / steps          # mean history over the steps
    step_gb = (weight_bytes + kv_bytes) / 1e9
    prefill_tflops = 76.4 * B_ / 1e3 / (ms_prefill * 1e-3)
    roofline = {"prefill": {"bound": "tensor", "ms": ms_prefill, "achieved": prefill_tflops, "peak": pk["tflops"], "unit": "TFLOP/s",
                            "frac": prefill_tflops / pk["tflops"], "algorithmic_gflop_per_answer": 76.4},
                "decode_step": {"bound": "hbm", "ms": ms_step, "steps": steps, "achieved": step_gb / (ms_step * 1e-3), "peak": pk["hbm"],
                                "unit": "GB/s", "frac": step_gb / (ms_step * 1e-3) / pk["hbm"], "gbytes_per_step": step_gb,
                                "weights_gb": weight_bytes / 1e9, "kv_history_gb": kv_bytes / 1e9,
                                "note": "a decode step is ~170 dependent launches (7 per layer): split-K projection GEMMs, KV-cache "
                                        "attention, residual + LayerNorm glue; bound by their latency chain, not by bytes (DESIGN.md)"}}
    return {"metric": "few_shot_vqa_answers_per_sec", "value": c["batch"] / (ms * 1e-3), "unit": "answers/s", "ms_per_batch": ms,
            "config": {"workload": "BASELINE configs[3]: 4-shot in-context prefixes, GPT-2 medium, batch 128, 10 new tokens, "
                                   "KV-cached greedy decode, host tensors in / token tensor out", "prompt_len": T0,
                       "tokens_out_shape": list(out.shape)},
            "algorithmic_gflop_per_answer": 82.8, "tflops": 82.8 * c["batch"] / 1e3 / (ms * 1e-3), "roofline": roofline}


def bench_rices(with_cpu=True):
    """RICES in-context example retrieval (the step before the few-shot path): faiss.normalize_L2 + IndexFlatIP.search
    (k = 2048) over a VQA2-train-sized database of CLIP text embeddings, 4096 queries per call."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import rices_bench
    r = rices_bench.run(M=4096, N=443757, D=768, k=2048, reps=10)
    if with_cpu:
        import numpy as np
        from oracle import rices as orc
        g = np.random.default_rng(0)
        base = g.standard_normal((1, 768)).astype(np.float32)
        db = (base + 0.5 * g.standard_normal((443757, 768))).astype(np.float32)
        q = (base + 0.5 * g.standard_normal((32, 768))).astype(np.float32)
        t0 = time.perf_counter()
        orc.knn_inner_product(q, db, 2048)
        sec = time.perf_counter() - t0
        r["cpu_port"] = {"value": 32 / sec, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": "32 queries against the same database (numpy float64, oracle/rices.py; includes normalising the database)"}
    return r


if __name__ == "__main__":
    main()
