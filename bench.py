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
           host memory inside the timed region; at N > 1 also the two collectives of SURVEY.md 8e -- all_gather of the
           frame counts, gather of the PCM rows to rank 0); wall clock bracketed by synchronisation.
N > 1    : one process per GPU (torchrun), utterances sharded over ranks, weights replicated, no data-path
           collective during decode; barrier + max over ranks; `value` = all ranks' frames / that time ("weak" scaling:
           8 utterances per GPU).  The line also carries `strong`: BASELINE.json configs[4], a GLOBAL batch of 32
           utterances split over the N ranks (32/N per GPU), measured the same way.
`by_batch`: the metric is quoted at batch 1 / 8 / 32: the batch-1 and batch-32 legs run (shorter) in the same process
           at N = 1 and are reported beside the batch-8 headline, each with its own roofline fraction.
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
    ap.add_argument("--cpu-frames", type=int, default=32, help="frames per utterance of a CPU-arm sample")
    ap.add_argument("--no-by-batch", action="store_true", help="skip the batch-1 / batch-32 legs (N = 1)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (N > 1)")
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


def ncu_traffic(kernel: str, model: str, batch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per FRAME of the decode kernel, from the committed `ncu --set full`
    capture of this kernel at this shape (profiles/traffic.json, written from the capture's raw page), or None: a
    number from a profiler run, not a measurement of this run (B200_PROFILING.md)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    for e in json.load(open(p)).get("captures", []):
        if e["kernel"] == kernel and e["model"] == model and e["batch"] == batch:
            return {"bytes_per_frame": e["dram_bytes_per_frame"], "source": e["source"]}
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_prompts(spec, n, first):
    from qwen3_tts_rs_b200 import weights as W
    return [W.synthetic_prompt(first + i, spec) for i in range(n)]


class CpuArm:
    """The reference's CPU path (F32), restated by the oracle (torch/MKL), on the host cores: one utterance at a time,
    as the reference runs (it has no batching).  A sample = prefill + `frames` decode frames + vocoder of ONE utterance of
    the GPU arm's prompt set (utterance index `utt`, seed 42 + utt)."""

    def __init__(self, spec, talker_w, vocoder_w, threads):
        import torch
        from oracle import model as OM, vocoder as OV
        torch.set_num_threads(threads)
        self.spec, self.threads = spec, threads
        self.tk, self.cp = OM.Talker(spec, talker_w, OM.F32P), OM.CodePredictor(spec, talker_w, OM.F32P)
        self.voc = OV.Vocoder(spec.vocoder, vocoder_w)

    def sample(self, utt, frames):
        from oracle import generate as OG, sampling as osmp
        from qwen3_tts_rs_b200 import spec as S
        ids = build_prompts(self.spec, 1, utt)[0]
        t0 = time.perf_counter()
        emb = self.tk.custom_voice_embeds(ids, S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"])
        fr = OG.prefill_and_generate(self.tk, self.cp, emb, ids, osmp.GenerationConfig(max_new_tokens=frames), 42 + utt)
        if fr:
            self.voc.decode(OG.codes_to_tensor(fr))
        return len(fr), time.perf_counter() - t0


def workload_config(spec, B, world, F, scaling="weak"):
    return {"workload": f"{spec.name} CustomVoice(ryan) non-streaming: prefill + {F} decode frames + vocoder, "
                        f"batch {B} per GPU (BASELINE.json configs[2]), synthetic prompt set of SURVEY.md 8d",
            "model": spec.name, "batch_per_gpu": B, "global_batch": B * world, "frames_per_step": F,
            "l2_policy": "inputs larger than L2 (weights 3.9 GB re-streamed every frame; no flush needed)",
            "vocoder_dtype": "f32"}


def run_reference(args, spec, rank, world):
    """--impl reference: rank 0 alone times the reference's CPU implementation of the path (oracle port, torch F32 on all
    host cores) on the GPU arm's config: every step is a bounded sample of that workload -- one utterance of the same
    prompt set (step s takes utterance s mod batch, seed 42 + that index), `--cpu-frames` decode frames instead of 256."""
    if rank != 0:
        return
    import torch
    from qwen3_tts_rs_b200 import weights as W
    cores = os.cpu_count() or 1
    tw = W.make_talker_weights(spec, dtype=torch.float32)
    vw = W.make_vocoder_weights(spec.vocoder)
    arm = CpuArm(spec, tw, vw, cores)
    F = args.cpu_frames
    for _ in range(max(1, min(args.warmup, 3))):
        arm.sample(0, 2)
    tot_f, tot_t = 0, 0.0
    for s in range(args.steps):
        f, t = arm.sample(s % args.batch, F)
        tot_f += f
        tot_t += t
    value = tot_f / tot_t
    sample = (f"per step: 1 utterance of the batch-{args.batch} prompt set x {F} frames (prefill + decode loop + vocoder), "
              f"run one at a time as the reference does (no batching), torch F32 (MKL) restatement of the reference CPU path")
    out = {
        "impl": "reference", "metric": "audio_frames_per_sec", "value": value, "unit": "frames/s",
        "rtf": (tot_t / (tot_f * 0.08)) if tot_f else None,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(spec, args.batch, max(1, args.gpus), args.frames),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


class Leg:
    """One workload on this rank's GPU: `B` utterances (global utterance indices first..first+B), F frames each."""

    def __init__(self, tts, spec, lib, B, F, first, local_rank):
        import torch
        from qwen3_tts_rs_b200 import api, shard
        self.tts, self.spec, self.B, self.F, self.first = tts, spec, B, F, first
        self.prompts = build_prompts(spec, B, first)
        self.seeds = shard.utterance_seeds(42, first, first + B)
        self.pp = [tts.custom_voice_prompt(t, "ryan", "english") for t in self.prompts]
        self.lmax = max(len(p[0]) for p in self.pp)
        self.sess = api.Session(tts.model, B, api.SynthesisOptions(max_length=F), self.seeds, max_seq=self.lmax + F + 8)
        self.trailing = [list(t[1:]) for t in self.prompts]
        self.stream = torch.cuda.ExternalStream(lib.q3_session_stream(self.sess.handle), device=torch.device("cuda", local_rank))

    def _prompt(self):
        s = self.sess
        s.reset(self.seeds)
        s.prefill_ids([p[0] for p in self.pp], [p[1] for p in self.pp])
        s.set_trailing_ids(self.trailing)

    def step_device(self):
        self._prompt()
        self.sess.generate_async(self.F)
        self.sess.vocode(self.F, to_host=False)

    def step_e2e(self):
        self._prompt()
        codes, n = self.sess.generate(self.F)
        pcm = self.sess.vocode(self.F, to_host=True)
        return codes, n, pcm

    def h2d_bytes(self):
        return int(sum(len(p[0]) * 8 for p in self.pp) + sum(len(t) * 4 for t in self.trailing))

    def close(self):
        self.sess.close()


def measure(leg, steps, warmup, rank, world, n_total, want_clocks=False, local_rank=0):
    """Times `steps` steps of one leg: device-resident (CUDA events on the session stream), the decode loop alone, and end
    to end through the host-buffer API (+ the two collectives at N > 1).  Max over ranks.  Returns a dict on every rank."""
    import torch
    import torch.distributed as dist
    from qwen3_tts_rs_b200 import lib as L, shard
    lib = L.load()
    sess, F = leg.sess, leg.F

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        leg.step_device()
    sess.synchronize()
    # ---- device-resident.  The working set of one step (3.9 GB of bf16 weights streamed every frame + 0.46 GB vocoder
    # weights + multi-GB vocoder activations) is far larger than the 126 MB L2, so no explicit L2 flush is needed.
    clocks = ClockSampler(local_rank) if want_clocks else None
    barrier()
    if clocks:
        clocks.start()
    launches0 = lib.q3_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dec_ms = 0.0
    ev0.record(leg.stream)
    for _ in range(steps):
        leg.step_device()
        dec_ms += sess.timing().decode_ms
    ev1.record(leg.stream)
    barrier()
    ms_dev = ev0.elapsed_time(ev1)
    launches = lib.q3_kernel_launch_count() - launches0
    clk = clocks.stop() if clocks else None
    # ---- the decode loop alone (CUDA events around the frame loop of one step)
    leg._prompt()
    sess.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(leg.stream)
    sess.generate_async(F)
    e1.record(leg.stream)
    sess.synchronize()
    loop_ms = e0.elapsed_time(e1)
    prefill_ms = sess.timing().prefill_ms
    _, nfr = sess.get_codes(F)
    frames_run = int(max(nfr)) if len(nfr) else F
    # ---- end to end: host buffers in and out, and (N > 1) the collectives of the sharded API inside the timed region
    def e2e_step():
        codes, n, pcm = leg.step_e2e()
        if world > 1:
            all_counts = shard.gather_frame_counts(n.tolist(), n_total, rank, world, device="cuda")
            rows = shard.gather_pcm(pcm, n.tolist(), all_counts.tolist(), rank, world, leg.spec.vocoder.total_upsample, 0, "cuda")
            assert (rows is None) == (rank != 0) and (rows is None or len(rows) == n_total)
        return codes, n, pcm

    e2e_step()        # untimed: NCCL connects the point-to-point channels of `gather` lazily, on its first call
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(steps):
        codes, n, pcm = e2e_step()
        d2h = int(codes.nbytes + n.nbytes + pcm.nbytes)
    barrier()
    e2e_s = time.perf_counter() - t0
    frames_step = int(n.sum())
    times = torch.tensor([ms_dev, e2e_s * 1e3, loop_ms], dtype=torch.float64, device="cuda")
    fr = torch.tensor([float(frames_step)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(fr, op=dist.ReduceOp.SUM)
    ms_dev_max, e2e_ms_max, loop_ms_max = [float(x) for x in times.tolist()]
    total = float(fr.item())
    return {"frames_per_step_all_ranks": total, "ms_dev": ms_dev_max, "e2e_ms": e2e_ms_max, "loop_ms": loop_ms_max,
            "loop_ms_local": loop_ms, "frames_run": frames_run, "prefill_ms": prefill_ms, "vocoder_ms": dec_ms / steps,
            "launches": int(launches), "clocks": clk, "h2d": leg.h2d_bytes(), "d2h": d2h, "steps": steps,
            "value": total * steps / (ms_dev_max / 1e3), "e2e_value": total * steps / (e2e_ms_max / 1e3)}


def roofline_of(spec, B, lmax, m, kernel, peak, peak_src):
    """Decode step (one frame for all B rows of this GPU) against the measured HBM copy bandwidth; algorithmic bytes =
    SURVEY.md 8d bytes_step(B, mean context length)."""
    from qwen3_tts_rs_b200 import spec as S
    ctx = lmax + m["frames_run"] / 2.0
    bytes_step = S.step_bytes(spec, B, ctx)
    t_frame = (m["loop_ms_local"] / 1e3) / max(1, m["frames_run"])
    achieved = bytes_step / t_frame / 1e9
    tr = ncu_traffic(kernel, spec.name, B)
    return {"kernel": f"{kernel}, per frame (one persistent cooperative launch runs 16 frames: 15 code-predictor passes + "
                      f"{spec.layers}-layer talker step + codec head + sampler for all rows)",
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
            "traffic": tr["bytes_per_frame"] if tr else None,
            "traffic_source": tr["source"] if tr else "no ncu --set full capture of this kernel at this shape is committed",
            "algorithmic_bytes_per_launch": bytes_step, "launch_ms": t_frame * 1e3}


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
    kernel = None      # set from the session once it exists (which generation of the persistent kernel actually runs)
    peak, peak_src = measured_peaks()
    B, F = args.batch, args.frames
    warm = max(args.warmup, 3)

    # ---- headline: weak scaling, B utterances per GPU ------------------------------------------------
    leg = Leg(tts, spec, lib, B, F, rank * B, local_rank)
    kernel = {0: "multi-kernel CUDA graph (gemv_kernel ...)", 1: "decode_frames_mega_kernel", 2: "decode_frames_mega2_kernel",
              3: "decode_frames_mega3_kernel", 4: "decode_frames_mega4_kernel", 5: "decode_frames_mega5_kernel"}[leg.sess.decode_generation()]
    m = measure(leg, args.steps, warm, rank, world, B * world, want_clocks=True, local_rank=local_rank)
    lmax = leg.lmax
    leg.close()

    # ---- batch 1 / 32 legs (N = 1) and the strong-scaling leg (N > 1: global batch 32 split over the ranks) -------
    by_batch, strong = None, None
    if world == 1 and not args.no_by_batch:
        by_batch = {}
        for b in (1, 32):
            if b == B:
                continue
            lg = Leg(tts, spec, lib, b, F, 0, local_rank)
            mb = measure(lg, 2, 3, rank, world, b)
            by_batch[str(b)] = (mb, lg.lmax)
            lg.close()
    if world > 1 and not args.no_strong and 32 % world == 0:
        bs = 32 // world
        lg = Leg(tts, spec, lib, bs, F, rank * bs, local_rank)
        strong = (measure(lg, max(2, min(args.steps, 5)), 3, rank, world, 32), lg.lmax, bs)
        lg.close()

    if rank == 0:
        def summary(mm, b, lm):
            total = mm["frames_per_step_all_ranks"]
            return {"batch_per_gpu": b, "value": mm["value"], "unit": "frames/s",
                    "rtf": (mm["ms_dev"] / 1e3) / (total * mm["steps"] * 0.08),
                    "e2e": mm["e2e_value"], "e2e_rtf": (mm["e2e_ms"] / 1e3) / (total * mm["steps"] * 0.08),
                    "ms_per_frame": mm["loop_ms"] / max(1, mm["frames_run"]), "vocoder_ms": mm["vocoder_ms"],
                    "steps": mm["steps"], "roofline_frac": roofline_of(spec, b, lm, mm, kernel, peak, peak_src)["frac"]}

        total = m["frames_per_step_all_ranks"]
        out = {
            "metric": "audio_frames_per_sec", "value": m["value"], "unit": "frames/s",
            "rtf": (m["ms_dev"] / 1e3) / (total * args.steps * 0.08),
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": m["ms_dev"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(spec, B, world, F),
            "breakdown_ms_per_step": {"decode_loop": m["loop_ms"], "frames_in_loop": m["frames_run"],
                                      "ms_per_frame": m["loop_ms"] / max(1, m["frames_run"]),
                                      "vocoder": m["vocoder_ms"], "prefill": m["prefill_ms"],
                                      "vocoder_tflops_f32_equiv": (total / world) * 4.959e9 / (m["vocoder_ms"] / 1e3) / 1e12},
            "roofline": roofline_of(spec, B, lmax, m, kernel, peak, peak_src),
            "e2e": {"value": m["e2e_value"], "unit": "frames/s",
                    "rtf": (m["e2e_ms"] / 1e3) / (total * args.steps * 0.08),
                    "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                    "collectives_in_timed_region": (["all_gather(frame counts)", "gather(pcm -> rank 0)"] if world > 1 else [])},
            "gpu_launches": m["launches"],
            "clocks": m["clocks"],
        }
        if by_batch is not None:
            out["by_batch"] = {str(B): summary(m, B, lmax)}
            for k, (mb, lm) in by_batch.items():
                out["by_batch"][k] = summary(mb, int(k), lm)
        if strong is not None:
            ms_, lm, bs = strong
            out["strong"] = dict(summary(ms_, bs, lm), scaling="strong", global_batch=32,
                                 workload="BASELINE.json configs[4]: 32 utterances sharded data-parallel over the ranks")
        if not args.no_cpu_baseline:
            # the reference's CPU path on this box's host cores: median of 3 bounded samples (utterances 0, 1, 2 of the same
            # prompt set, --cpu-frames frames each), after a warm-up sample
            cores = os.cpu_count() or 1
            arm = CpuArm(spec, {k: v.float() for k, v in tw.items()}, vw, cores)
            arm.sample(0, 2)
            rates = []
            for u in range(3):
                f, t = arm.sample(u % B, max(2, args.cpu_frames // 2))
                rates.append(f / t)
            rates.sort()
            med = rates[1]
            out["cpu_baseline"] = {"value": med, "unit": "frames/s", "cores": cores, "kind": "port",
                                   "rtf": 1.0 / (med * 0.08), "samples": rates,
                                   "sample": f"median of 3 samples, each 1 utterance x {max(2, args.cpu_frames // 2)} frames (prefill + decode loop "
                                             "+ vocoder), batch 1, torch F32 (MKL) restatement of the reference CPU path"}
            b1 = out.get("by_batch", {}).get("1")
            if b1:
                # per-stream comparison: one utterance at a time on both sides
                out["cpu_baseline"]["gpu_batch1_e2e_over_cpu"] = b1["e2e"] / med
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
