import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo')
from neuralnj_b200 import _lib
L = _lib.lib()
A = torch.zeros(128, 64, device="cuda"); B = torch.zeros(64, 256, device="cuda"); D = torch.zeros(128, 256, device="cuda")
names = {0: "N=64", 1: "N=128", 2: "N=256", 3: "N=32"}
for two in (0, 16):
    for bmn in (0, 8):
        for ta in (0, 4):
            for n in (3, 0, 1, 2):
                v = n | ta | bmn | two
                if n == 2 and two: continue
                _lib.check(L.nnj_tc_selftest(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), 2000 + v, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
                torch.cuda.synchronize()
                print(f"{names[n]:6s} A={'TMEM' if ta else 'smem'} B={'MN' if bmn else 'K '}-major acc={'2 alternating' if two else '1'}: {float(D[0,0]):7.1f} clk/UMMA (issue loop {float(D[0,1]):6.1f})")
