"""Debug timeline of CTA 0 of lora_gemm_kernel (clock64 stamps; see SDT_TRACE in csrc/lora_gemm.cu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scal_sdt_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
NAMES = {0: "entry", 1: "prologue done", 2: "first TMA issued", 62: "epi: stores drained", 63: "all warps done", 64: "dealloc done"}
for t in range(6):
    NAMES[8 + 4 * t] = f"mma t{t}: acc_empty ok"
    NAMES[9 + 4 * t] = f"mma t{t}: first k-block landed"
    NAMES[10 + 4 * t] = f"mma t{t}: K loop issued"
    NAMES[11 + 4 * t] = f"mma t{t}: tail issued"
    NAMES[40 + t] = f"side t{t}: tail operands ready"
    NAMES[48 + 2 * t] = f"epi t{t}: acc_full"
    NAMES[49 + 2 * t] = f"epi t{t}: accumulator released"


for t in range(6):
    NAMES[70 + t] = f"epi t{t}: stores issued"
    NAMES[80 + t] = f"epi t{t}: geglu pair barrier passed"
    NAMES[90 + t] = f"epi t{t}: geglu act + proj stores issued"
for cbi in range(0):
    for h in range(2):
        NAMES[70 + 8 * cbi + 4 * h] = f"  epi t0 w6 blk{cbi} sub{h}: ld issued"
        NAMES[71 + 8 * cbi + 4 * h] = f"  epi t0 w6 blk{cbi} sub{h}: ld done"
        NAMES[72 + 8 * cbi + 4 * h] = f"  epi t0 w6 blk{cbi} sub{h}: staged"
    NAMES[73 + 8 * cbi] = f"  epi t0 w6 blk{cbi}: written"


def run_geglu(M, K, I, R):
    x = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn(2 * I, K, device=dev).bfloat16()
    b = torch.randn(2 * I, device=dev)
    A = torch.randn(R, K, device=dev).bfloat16()
    B = torch.randn(2 * I, R, device=dev).bfloat16()
    proj = torch.empty(M, 2 * I, device=dev, dtype=torch.bfloat16)
    act = torch.empty(M, I, device=dev, dtype=torch.bfloat16)
    t = torch.empty(M, R, device=dev, dtype=torch.bfloat16)
    trace = torch.zeros(128, dtype=torch.int64, device=dev)

    def call():
        _lib.check(lib.sdt_lora_linear_geglu_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), A.data_ptr(), B.data_ptr(), 0.5, proj.data_ptr(),
                                                 act.data_ptr(), t.data_ptr(), M, K, I, R, 1, torch.cuda.current_stream().cuda_stream))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    lib.sdt_debug_set(10, trace.data_ptr())
    call()
    torch.cuda.synchronize()
    lib.sdt_debug_set(10, 0)
    tr = trace.cpu().tolist()
    t0 = tr[0]
    print(f"--- GEGLU epilogue M={M} K={K} I={I} R={R}: {e0.elapsed_time(e1) * 100:.1f} us per launch (cycles since entry, CTA 0)")
    for slot, v in sorted(((s, v) for s, v in enumerate(tr) if v), key=lambda kv: kv[1]):
        print(f"{v - t0:8d}  {NAMES.get(slot, slot)}")


def run(M, K, N, R, bias=True):
    x = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    A = torch.randn(R, K, device=dev).bfloat16()
    B = torch.randn(N, R, device=dev).bfloat16()
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t = torch.empty(M, R, device=dev, dtype=torch.bfloat16)
    trace = torch.zeros(128, dtype=torch.int64, device=dev)

    def call():
        _lib.check(lib.sdt_lora_linear_fwd(x.data_ptr(), w.data_ptr(), _lib.ptr(b), A.data_ptr(), B.data_ptr(), 0.5, y.data_ptr(),
                                           t.data_ptr(), M, K, N, R, 1, torch.cuda.current_stream().cuda_stream))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    lib.sdt_debug_set(10, trace.data_ptr())
    call()
    torch.cuda.synchronize()
    lib.sdt_debug_set(10, 0)
    tr = trace.cpu().tolist()
    t0 = tr[0]
    print(f"--- M={M} K={K} N={N} R={R} bias={bias} (cycles since entry, CTA 0)")
    for slot, v in sorted(((s, v) for s, v in enumerate(tr) if v), key=lambda kv: kv[1]):
        print(f"{v - t0:8d}  {NAMES.get(slot, slot)}")


mode = sys.argv[1] if len(sys.argv) > 1 else "single"
if mode == "geglu":
    for shape in [(32768, 320, 1280, 16), (8192, 640, 2560, 16)]:
        run_geglu(*shape)
    sys.exit(0)
if mode == "single":
    lib.sdt_debug_set(11, 1)          # force the single-CTA kernel
else:
    lib.sdt_debug_set(14, 64)         # CTA-pair kernel for every K
print("==== kernel:", mode)
if len(sys.argv) > 2 and sys.argv[2] == "gs":
    for gs in (2, 1):
        lib.sdt_debug_set(20, gs)
        print("==== column tiles per work item:", gs)
        run(32768, 2560, 320, 16)
    lib.sdt_debug_set(20, 0)
    sys.exit(0)
for shape in [(32768, 320, 320, 16), (8192, 640, 640, 16), (2048, 1280, 1280, 16), (32768, 320, 2560, 16), (8192, 640, 5120, 16)]:
    run(*shape)
