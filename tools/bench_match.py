"""Matcher-only timing (CUDA events) at the BASELINE shape, every mode interleaved in one process.
ITERS launches per measurement (default 50); LONG=seconds adds a sustained loop per mode with nvidia-smi clocks / power
sampled every 100 ms (what the kernel does under the power cap)."""
import sys, os, subprocess, threading, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gadm_b200  # noqa
from gadm_b200 import ops, synth, _lib
from gadm_b200._lib import MATCH_MODES

dev = torch.device("cuda", 0)
B, N, M, D = int(os.environ.get("B", "8")), int(os.environ.get("N", "12800")), 8192, int(os.environ.get("D", "128"))
ITERS = int(os.environ.get("ITERS", "50"))
LONG = float(os.environ.get("LONG", "0"))
regime = os.environ.get("REGIME", "planted")
rgbd, mesh, _ = synth.descriptors(B, N, M, D, n_obj=8, regime=regime, seed=2000)
xyz = synth.model_bank_xyz(8, M).to(dev)
obj = torch.arange(B, dtype=torch.int32, device=dev) % 8
cols, aux = ops.prep_model(mesh.to(dev), xyz, 0)
cols_n, aux_n = ops.prep_model(mesh.to(dev), xyz, 2)
rows, rinv, pad = ops.prep_rows(rgbd.to(dev), 0, 0)
flop = 2.0 * N * M * D * B
PEAK = 1658.8

# (label, mode, bf16n operands, config switches)
CASES = [("argmax", "argmax", False, {}), ("argmax_pair", "argmax", False, {"match.alt": 0, "match.pair": 1}),
         ("argmax_pair_cta2", "argmax", False, {"match.alt": 0, "match.pair": 1, "match.cta2": 1}),
         ("argmax_rt2", "argmax", False, {"match.alt": 0, "match.pair": 0, "match.rt": 2}),
         ("argmax_single", "argmax", False, {"match.alt_cta2": 0}), ("soft", "soft", False, {}),
         ("argmax_bf16n_single", "argmax_bf16n", True, {"match.alt_cta2": 0}),
         ("argmax_unit_single", "argmax_unit", True, {"match.alt_cta2": 0}), ("soft_cta2", "soft", False, {"match.cta2": 1}),
         ("argmax_bf16n", "argmax_bf16n", True, {}), ("argmax_unit", "argmax_unit", True, {})]
if os.environ.get("CASES"):
    CASES = [c for c in CASES if c[0] in os.environ["CASES"].split(",")]


def launch(mode, n, bf16n):
    c, a = (cols_n, aux_n) if bf16n else (cols, aux)
    for _ in range(n):
        ops.match_fwd(rows, rinv, pad, c, a, None, obj, 16.0, 0, MATCH_MODES[mode])


def timed(mode, n, bf16n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    launch(mode, n, bf16n)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


class Smi:
    def __init__(self):
        self.lines = []
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                                   "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._rd, daemon=True).start()

    def _rd(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        self.p.terminate()
        clk, pw, cap = [], [], 0
        for ln in self.lines:
            q = [x.strip() for x in ln.split(",")]
            try:
                clk.append(float(q[0])); pw.append(float(q[1])); cap += q[2].lower().startswith("active")
            except Exception:
                pass
        return clk, pw, cap


for rep in range(int(os.environ.get("REPS", "2"))):
    for label, mode, bf16n, sw in CASES:
        for k, v in sw.items():
            _lib.config_set(k, v)
        launch(mode, 3, bf16n)
        ms = timed(mode, ITERS, bf16n)
        for k in sw:
            _lib.config_set(k, -1)
        print(f"{label:18s} B={B} N={N} d={D} {regime}: {ms:.4f} ms  {flop/ms/1e9:7.1f} TFLOP/s  frac {flop/ms/1e9/PEAK:.3f}", flush=True)

if LONG > 0:
    for label, mode, bf16n, sw in CASES:
        for k, v in sw.items():
            _lib.config_set(k, v)
        launch(mode, 3, bf16n)
        torch.cuda.synchronize()
        smi = Smi()
        time.sleep(0.3)
        t0, chunks = time.time(), []
        while time.time() - t0 < LONG:
            chunks.append(timed(mode, 200, bf16n))
        clk, pw, cap = smi.stop()
        for k in sw:
            _lib.config_set(k, -1)
        half = chunks[len(chunks) // 2:]
        ms = statistics.mean(half)
        busy = [(c, p) for c, p in zip(clk, pw) if p > 200]
        cm = statistics.median([c for c, _ in busy]) if busy else float("nan")
        pm = statistics.median([p for _, p in busy]) if busy else float("nan")
        print(f"sustained {label:18s}: {ms:.4f} ms  frac {flop/ms/1e9/PEAK:.3f} (first chunk {chunks[0]:.4f} ms)  "
              f"sm clock median {cm:.0f} MHz, power median {pm:.0f} W, sw_power_cap samples {cap}/{len(clk)}", flush=True)
