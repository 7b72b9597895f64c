"""Kernel-only time of the tcgen05 GEMM at the encoder's shapes (operands pre-packed by the plane cache: versions not bumped)."""
import os, sys, numpy as np, ctypes as C
sys.path.insert(0, os.getcwd())
import pydynet_b200 as pdn
from pydynet_b200.backend import lib
rng = np.random.default_rng(0)
PEAK = 1416e12
SHAPES = {'encoder': [(65536, 512, 512), (65536, 1536, 512), (65536, 512, 1536), (512, 512, 65536), (512, 1536, 65536), (8192, 8192, 8192)],
          'other': [(256, 500, 2450), (2450, 500, 256), (256, 2450, 500), (262144, 1536, 512), (512, 1536, 262144), (512, 2048, 262144), (1024, 1024, 1024), (2048, 2048, 2048), (4096, 512, 4096), (512, 4096, 4096)]}
for (M, N, K) in SHAPES[os.environ.get('SHAPES', 'encoder')]:
    with pdn.Device("cuda:0"):
        a = pdn.backend.array(rng.standard_normal((M, K)).astype(np.float32)); b = pdn.backend.array(rng.standard_normal((K, N)).astype(np.float32))
        out = pdn.backend.empty((M, N), np.float32)
        for _ in range(3):
            pdn.backend.gemm_into(out, a, b)
        pdn.cuda.synchronize()
        e0, e1 = C.c_void_p(), C.c_void_p()
        lib.call("pdn_event_create", C.byref(e0)); lib.call("pdn_event_create", C.byref(e1))
        lib.call("pdn_event_record", e0)
        for _ in range(10):
            pdn.backend.gemm_into(out, a, b)
        lib.call("pdn_event_record", e1)
        ms = C.c_float(); lib.load().pdn_event_elapsed_ms(e0, e1, C.byref(ms))
        t = ms.value / 10 / 1e3
        print(f"M{M} N{N} K{K}: {t * 1e6:.1f} us, {2.0 * M * N * K / t / 1e12:.0f} TFLOP/s alg, {3 * 2.0 * M * N * K / t / PEAK:.2f} of BF16x3 ceiling", flush=True)
