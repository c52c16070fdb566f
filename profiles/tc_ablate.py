"""Ablation timing of the tensor-core edge kernel at BASELINE configs[2] size: which role bounds the tile time?
flags: 1 skip weight copies, 2 skip producer work, 4 skip epilogue work, 8 skip MMAs (results are garbage when set)."""
import ctypes as C
import os
import sys

os.environ.setdefault('DSB_INSTRUMENT', '1')     # the product library compiles the switches out

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffsbdd_b200 import _native, synthetic as syn  # noqa: E402
from diffsbdd_b200.config import FULLATOM_COND  # noqa: E402
from diffsbdd_b200.dynamics import EGNNDynamics  # noqa: E402

B = 64
cfg = FULLATOM_COND
net = EGNNDynamics.from_config(cfg, device='cuda')
net.load_state_dict(syn.synthetic_state_dict(cfg, 0))
net.eval()
net.defer_status_check = True
inp = [x.cuda() for x in syn.synthetic_denoiser_inputs(cfg, [25] * B, [175] * B, seed=3)]
lib = _native.load()
lib.dsb_debug_set_tc_flags.argtypes = [C.c_int]
lib.dsb_debug_read_tc_prof.argtypes = [C.POINTER(C.c_uint64)]


def read_prof():
    buf = (C.c_uint64 * 64)()
    lib.dsb_debug_read_tc_prof(buf)
    return list(buf)
with torch.no_grad():
    net(*inp)
    net.set_profiling(True)
    for flags in [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 8, 3, 5, 6, 7, 15]:
        lib.dsb_debug_set_tc_flags(flags)
        read_prof()
        net(*inp); net(*inp)
        net.collect_profile(reset=True)
        for _ in range(5):
            net(*inp)
        p = net.collect_profile(reset=True)
        if flags & 512:
            c = read_prof()
            nt, npv = max(c[3], 1), max(c[13], 1)
            print(f'   epilogue warp0 per virtual tile (cycles): wait {c[0] / nt:8.0f}  pass1 {c[1] / nt:8.0f}  pass2 {c[2] / nt:8.0f}   [TMEM ld: pass1 {c[4] / nt:7.0f} pass2 {c[5] / nt:7.0f}; pass2 scale+STS+syncwarp {c[6] / nt:7.0f}; wait::st {c[28] / nt:7.0f}; chunk sums+RED {c[29] / nt:7.0f}]  (n={nt})')
            print(f'   producer thread0 per virtual tile (cycles): tile-start {c[8] / npv:7.0f}  compute+gather {c[9] / npv:8.0f}  wait-empty {c[10] / npv:8.0f}  '
                  f'store {c[11] / npv:7.0f}  fence+arrive {c[12] / npv:7.0f}   (n={npv})')
        if flags & 512:
            print(f'   whole tile loop per CTA (cycles): epilogue warp0 {c[14] / max(c[15], 1):9.0f}   producer thread0 {c[24] / max(c[25], 1):9.0f}   '
                  f'(vtiles/CTA {nt / max(c[15], 1):.2f})')
            for tag, name in enumerate(('node GEMM', 'GCL', 'coord')):
                o = 32 + 8 * tag
                nc = max(c[o + 4], 1)
                print(f'   MMA thread, {name}: per chunk (cycles): wait-accumulator {c[o] / nc:7.0f}  wait-W {c[o + 1] / nc:7.0f}  wait-X {c[o + 2] / nc:7.0f}  issue+commit {c[o + 3] / nc:7.0f}   (chunks={nc})')
            ng = max(c[23], 1)
            print(f'   node GEMM CTA0 per launch (cycles): setup {c[16] / ng:7.0f}  epilogue-wait {c[17] / ng:8.0f}  epilogue-work {c[18] / ng:8.0f}  '
                  f'producers {c[19] / ng:8.0f}  body {c[20] / ng:8.0f}  teardown {c[21] / ng:7.0f}  tiles/launch {c[22] / ng:.2f}  (n={ng})')
        if flags & 512 and c[61]:
            ni = c[61]
            for k, name in enumerate(('phase 1 (K=2H)', 'phase 2', 'phase 3 (column tiles)')):
                o = 48 + 4 * k
                print(f'   node block, leader MMA thread per item (cycles), {name}: wait-A {c[o] / ni:8.0f}  wait-W {c[o + 1] / ni:8.0f}  wait-peer-W {c[o + 2] / ni:8.0f}  issue {c[o + 3] / ni:8.0f}')
            print(f'   node block per item: accumulator waits {c[60] / ni:8.0f}  whole loop {c[62] / ni:9.0f}   (items={ni})')
        print(f'flags={flags:2d}  edge_gcl {p["edge_gcl"]["ms"] / 30 * 1e3:8.1f} us/launch   edge_coord {p["edge_coord"]["ms"] / 30 * 1e3:8.1f}'
              f'   node_gemm {p["node_gemm"]["ms"] / 120 * 1e3:7.1f} us/launch', flush=True)
    lib.dsb_debug_set_tc_flags(0)
print('edges', int(net._status[1]))
