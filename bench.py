#!/usr/bin/env python
"""Benchmark of the Qwen3-TTS decode hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU (F32) path, oracle port

Metric (BASELINE.json): audio frames per wall-second (12.5 Hz codec frames, the reference's "Tok/s",
benches/e2e_bench.rs:334-344) with RTF = wall / audio seconds reported beside it, on the 1.7B CustomVoice
model, batch 8 per GPU, non-streaming (BASELINE.json configs[2]).  One STEP = one batch of utterances taken
through the whole hot path: prompt assembly + prefill -> F decode frames (code predictor, talker step,
sampling, EOS bookkeeping) -> vocoder over the F frames of every row.  Weights are synthetic (seeded, no
checkpoints exist offline); prompts are the synthetic prompt set of SURVEY.md §8d.

`value`  : frames/s with everything resident in HBM (codes and PCM stay on the device), CUDA-event timed.
`e2e`    : the same step through the host-buffer API (ids in from host memory, codes + PCM copied back to
           host memory inside the timed region); wall clock bracketed by synchronisation.
N > 1    : one process per GPU (torchrun), utterances sharded over ranks, weights replicated, no data-path
           collective; barrier + max over ranks; `value` = all ranks' frames / that time ("weak" scaling:
           8 utterances per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="1.7b")
    ap.add_argument("--batch", type=int, default=8, help="utterances per GPU")
    ap.add_argument("--frames", type=int, default=256, help="decode frames per utterance per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=6)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 8:
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        mx = None
        for r in self.rows:
            if len(r) >= 8 and r[2].replace(".", "").isdigit():
                mx = int(float(r[2]))
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# one `ncu --set full` capture of decode_frames_mega2_kernel at the bench shape (1.7B, batch 8): 73.30 GB read +
# 0.42 GB written per 16-frame launch (profiles/r1_mega2_full.summary.txt)
NCU_TRAFFIC_PER_FRAME = (73.301614e9 + 0.415258e9) / 16.0


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_prompts(spec, n, first):
    from qwen3_tts_rs_b200 import weights as W
    return [W.synthetic_prompt(first + i, spec) for i in range(n)]


def cpu_reference_run(spec, talker_w, vocoder_w, n_utt, frames, first_prompt=0, threads=None):
    """The reference's CPU path (F32), restated by the oracle, on the host cores.  Returns (frames, seconds)."""
    import torch
    from oracle import generate as OG, model as OM, sampling as osmp, vocoder as OV
    from qwen3_tts_rs_b200 import spec as S
    if threads:
        torch.set_num_threads(threads)
    tk, cp = OM.Talker(spec, talker_w, OM.F32P), OM.CodePredictor(spec, talker_w, OM.F32P)
    voc = OV.Vocoder(spec.vocoder, vocoder_w)
    cfg = osmp.GenerationConfig(max_new_tokens=frames)
    prompts = build_prompts(spec, n_utt, first_prompt)
    total, t0 = 0, time.perf_counter()
    for i, ids in enumerate(prompts):
        emb = tk.custom_voice_embeds(ids, S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"])
        fr = OG.prefill_and_generate(tk, cp, emb, ids, cfg, 42 + first_prompt + i)
        if fr:
            voc.decode(OG.codes_to_tensor(fr))
        total += len(fr)
    return total, time.perf_counter() - t0


def run_reference(args, spec, rank, world):
    """--impl reference: rank 0 alone times the reference's CPU implementation of the path (oracle port,
    torch F32 on all host cores) on a bounded sample per step."""
    if rank != 0:
        return
    import torch
    from qwen3_tts_rs_b200 import weights as W
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tw = W.make_talker_weights(spec, dtype=torch.float32)
    vw = W.make_vocoder_weights(spec.vocoder)
    frames_per_step = args.cpu_frames
    for _ in range(args.warmup):
        cpu_reference_run(spec, tw, vw, 1, 2)
    tot_f, tot_t = 0, 0.0
    for s in range(args.steps):
        f, t = cpu_reference_run(spec, tw, vw, 1, frames_per_step, first_prompt=s)
        tot_f += f
        tot_t += t
    value = tot_f / tot_t
    sample = f"1 utterance x {frames_per_step} frames per step (prefill + decode loop + vocoder), batch 1, F32"
    out = {
        "impl": "reference", "metric": "audio_frames_per_sec", "value": value, "unit": "frames/s",
        "rtf": (tot_t / (tot_f * 0.08)) if tot_f else None,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{spec.name} CustomVoice(ryan) non-streaming, reference CPU path (oracle port, torch F32/MKL)",
                   "model": spec.name, "batch": 1, "frames_per_step": frames_per_step},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from qwen3_tts_rs_b200 import spec as S
    spec = S.SPECS[args.model]

    if args.impl == "reference":
        run_reference(args, spec, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from qwen3_tts_rs_b200 import api, lib as L, weights as W

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- model (replicated per GPU) --------------------------------------------------------------
    tw = W.make_talker_weights(spec)
    vw = W.make_vocoder_weights(spec.vocoder)
    tts = api.Qwen3TTS.from_weights(spec, tw, vw, device=local_rank)
    lib = L.load()

    from qwen3_tts_rs_b200 import shard
    B, F = args.batch, args.frames
    first, last = shard.shard_range(B * world, rank, world)       # weak scaling: B utterances per GPU
    prompts = build_prompts(spec, B, first)
    seeds = shard.utterance_seeds(42, first, last)
    opts = api.SynthesisOptions(max_length=F)
    pp = [tts.custom_voice_prompt(t, "ryan", "english") for t in prompts]
    lmax = max(len(p[0]) for p in pp)
    sess = api.Session(tts.model, B, opts, seeds, max_seq=lmax + F + 8)
    trailing = [list(t[1:]) for t in prompts]
    stream = torch.cuda.ExternalStream(lib.q3_session_stream(sess.handle), device=torch.device("cuda", local_rank))
    up = spec.vocoder.total_upsample

    def step_device():
        sess.reset(seeds)
        sess.prefill_ids([p[0] for p in pp], [p[1] for p in pp])
        sess.set_trailing_ids(trailing)
        sess.generate_async(F)
        sess.vocode(F, to_host=False)

    def step_e2e():
        sess.reset(seeds)
        sess.prefill_ids([p[0] for p in pp], [p[1] for p in pp])
        sess.set_trailing_ids(trailing)
        codes, n = sess.generate(F)
        pcm = sess.vocode(F, to_host=True)
        return codes, n, pcm

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    sess.synchronize()

    # ---- timed: device-resident ---------------------------------------------------------------------
    # The working set of one step (3.9 GB of bf16 weights streamed ~16x per frame + 0.46 GB vocoder weights
    # + multi-GB vocoder activations) is far larger than the 126 MB L2, so no explicit L2 flush is needed.
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    launches0 = lib.q3_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gen_ms = dec_ms = 0.0
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
        t = sess.timing()
        gen_ms += t.generation_ms if t.generation_ms else 0.0
        dec_ms += t.decode_ms
    ev1.record(stream)
    barrier()
    ms_dev = ev0.elapsed_time(ev1)
    launches = lib.q3_kernel_launch_count() - launches0
    clk = clocks.stop()

    # decode-loop-only timing (CUDA events around the frame loop of one step)
    sess.reset(seeds)
    sess.prefill_ids([p[0] for p in pp], [p[1] for p in pp])
    sess.set_trailing_ids(trailing)
    sess.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sess.generate_async(F)
    e1.record(stream)
    sess.synchronize()
    loop_ms = e0.elapsed_time(e1)
    prefill_ms = sess.timing().prefill_ms
    _, nfr = sess.get_codes(F)
    frames_run = int(max(nfr)) if len(nfr) else F

    # ---- timed: end to end through the host-buffer API ----------------------------------------------------
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        codes, n, pcm = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    frames_step = int(n.sum())
    # the only collective of the path: per-utterance frame counts gathered to every rank
    all_counts = shard.gather_frame_counts(n.tolist(), B * world, rank, world, device="cuda")
    assert int(all_counts.sum()) >= frames_step

    times = torch.tensor([ms_dev, e2e_s * 1e3, loop_ms], dtype=torch.float64, device="cuda")
    fr = torch.tensor([float(frames_step)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(fr, op=dist.ReduceOp.SUM)
    ms_dev_max, e2e_ms_max, loop_ms_max = [float(x) for x in times.tolist()]
    total_frames_step = float(fr.item())

    if rank == 0:
        value = total_frames_step * args.steps / (ms_dev_max / 1e3)
        e2e_value = total_frames_step * args.steps / (e2e_ms_max / 1e3)
        peak, peak_src = measured_peaks()
        # roofline of the decode step (one replay of the frame graph = one frame for all B rows):
        # algorithmic bytes = SURVEY.md §8d bytes_step(B, mean context length)
        ctx = lmax + frames_run / 2.0
        bytes_step = S.step_bytes(spec, B, ctx)
        t_frame = (loop_ms / 1e3) / max(1, frames_run)
        achieved = bytes_step / t_frame / 1e9
        out = {
            "metric": "audio_frames_per_sec", "value": value, "unit": "frames/s",
            "rtf": (ms_dev_max / 1e3) / (total_frames_step * args.steps * 0.08),
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{spec.name} CustomVoice(ryan) non-streaming: prefill + {F} decode frames + vocoder, "
                                   f"batch {B} per GPU (BASELINE.json configs[2])",
                       "model": spec.name, "batch_per_gpu": B, "global_batch": B * world, "frames_per_step": F,
                       "l2_policy": "inputs larger than L2 (weights 3.9 GB re-streamed every frame; no flush needed)",
                       "vocoder_dtype": "f32"},
            "breakdown_ms_per_step": {"decode_loop": loop_ms_max, "frames_in_loop": frames_run,
                                      "ms_per_frame": loop_ms_max / max(1, frames_run),
                                      "vocoder": dec_ms / args.steps, "prefill": prefill_ms,
                                      "vocoder_tflops_f32_equiv": (total_frames_step / world) * 4.959e9 / (dec_ms / args.steps / 1e3) / 1e12},
            "roofline": {"kernel": "decode_frames_mega2_kernel, per frame (one persistent cooperative launch runs 16 frames: 15 code-predictor "
                                   "passes + 28-layer talker step + codec head + sampler for all rows)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src,
                         # dram__bytes_read + dram__bytes_write of one 16-frame launch / 16 (profiles/r1_mega2_full.summary.txt);
                         # below the algorithmic bytes because part of the code-predictor weights stay in L2 between passes
                         "traffic": NCU_TRAFFIC_PER_FRAME if (spec.name == "1.7b" and B == 8) else None,
                         "algorithmic_bytes_per_launch": bytes_step, "launch_ms": t_frame * 1e3},
            "e2e": {"value": e2e_value, "unit": "frames/s",
                    "rtf": (e2e_ms_max / 1e3) / (total_frames_step * args.steps * 0.08),
                    "h2d_bytes_per_step": int(sum(len(p[0]) * 8 for p in pp) + sum(len(t) * 4 for t in trailing)),
                    "d2h_bytes_per_step": int(codes.nbytes + n.nbytes + pcm.nbytes)},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        if not args.no_cpu_baseline:
            import torch as _t
            cores = os.cpu_count() or 1
            tw32 = {k: v.float() for k, v in tw.items()}
            cpu_reference_run(spec, tw32, vw, 1, 1, threads=cores)          # warm-up
            f, t = cpu_reference_run(spec, tw32, vw, 1, args.cpu_frames, threads=cores)
            out["cpu_baseline"] = {"value": f / t, "unit": "frames/s", "cores": cores, "kind": "port",
                                   "rtf": t / (f * 0.08) if f else None,
                                   "sample": f"1 utterance x {args.cpu_frames} frames (prefill + decode loop + vocoder), batch 1, "
                                             "torch F32 (MKL) restatement of the reference CPU path"}
        print(json.dumps(out), flush=True)
    sess.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
