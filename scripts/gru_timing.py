#!/usr/bin/env python
"""Where does a step of the bf16 GRU recurrence spend its cycles?  Builds csrc/gru_tc.cu alone with -DTAG_GRU_TIMING (clock64
stamps of one thread over 64 steps) and prints the mean cycles between the stamps.  Usage: python scripts/gru_timing.py"""
import ctypes
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csrc = os.path.join(ROOT, "texttoaudiogrounding_b200", "csrc")
so = "/tmp/libgru_timing.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                       "-DTAG_GRU_TIMING", *[f"-D{d}" for d in os.environ.get("GRU_EXP", "").split()], "-shared", "-I", csrc, "-I", os.path.join(ROOT, "include"),
                       os.path.join(csrc, "gru_tc.cu"), "-o", so])
lib = ctypes.CDLL(so)
B, T = 64, 250
gi = torch.randn(B, T, 1536, device="cuda")
w_hh = torch.randn(2, 768, 256, device="cuda") * 0.05
b_hh = torch.randn(1536, device="cuda") * 0.1
out = torch.empty(B, T, 512, device="cuda")
gates = torch.empty(B, T, 2, 4, 256, device="cuda")
P = ctypes.c_void_p
def stamps():
    buf = (ctypes.c_longlong * (64 * 8))()
    assert lib.tag_gru_timing_read(buf) == 0
    return [[buf[i * 8 + j] for j in range(8)] for i in range(64)]


def report(rows, names):
    if rows[0][0] == 0:
        print("  (no stamps)")
        return
    for j in range(6):
        print(f"  {names[j]:58s} {sum(r[j + 1] - r[j] for r in rows) / len(rows):8.1f} cycles")
    print(f"  {names[6]:58s} {sum(rows[i + 1][0] - rows[i][6] for i in range(63)) / 63:8.1f} cycles")
    print(f"  step total {sum(rows[i + 1][0] - rows[i][0] for i in range(63)) / 63:8.1f} cycles")
    if rows[2][7]:
        print(f"  prologue (weights -> registers, barrier init, cluster.sync) {rows[0][7] - rows[1][7]} cycles; "
              f"whole loop {rows[2][7] - rows[0][7]} cycles = {(rows[2][7] - rows[0][7]) / T:.1f} / step")


def timed(fn, label):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # keep the SMs busy first so that the clock is up when the (power-light) recurrence runs
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    for _ in range(20):
        a @ a
    s.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    print(f"{label}: {s.elapsed_time(e) * 1e3:.1f} us for {T} steps = {s.elapsed_time(e) * 1e3 / T:.3f} us / step")


stream = P(torch.cuda.current_stream().cuda_stream)
timed(lambda: lib.tag_gru_fwd_bf16(P(gi.data_ptr()), P(w_hh.data_ptr()), P(b_hh.data_ptr()), P(out.data_ptr()),
                                   P(gates.data_ptr()), B, T, stream), "forward")
names = ["loop top -> after mbarrier wait", "wait -> MMAs + partial stores issued", "-> __syncthreads done",
         "-> partial sums read", "-> gates + staging (st.async about to issue)", "-> st.async issued",
         "-> next loop top (HBM stores, bookkeeping)"]
report(stamps(), names)

d_out = torch.randn(B, T, 512, device="cuda")
dgi = torch.empty(B * T, 1536, device="cuda", dtype=torch.bfloat16)
dgh = torch.empty(2, B * T, 768, device="cuda", dtype=torch.bfloat16)
hprev = torch.empty(2, B * T, 256, device="cuda", dtype=torch.bfloat16)
timed(lambda: lib.tag_gru_bwd_bf16(P(d_out.data_ptr()), P(out.data_ptr()), P(gates.data_ptr()), P(w_hh.data_ptr()),
                                   P(dgi.data_ptr()), P(dgh.data_ptr()), P(hprev.data_ptr()), B, T, stream), "backward")
if os.environ.get("TAG_B200_GRU_OPERAND_EXCHANGE"):
    report(stamps(), ["loop top -> gate gradients", "-> staging + 3 st.async issued",
                      "-> HBM stores + prefetch of the next step",
                      "-> (loop exit test)", "-> after mbarrier wait", "-> 6 B loads + 12 MMAs + 8 partial stores",
                      "-> __syncthreads + partial sums -> next loop top"])
