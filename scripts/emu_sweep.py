"""Extended CPU sweep of the frame search's device source through the host emulation (tests/host_emu): DIA / HEX / UMH / STAR x
subme 0..7 x 8/10-bit (+ FULL), two CTUs each, every 2Nx2N PU against the reference's MotionEstimate -- for the shipped build and
with the shared vertical cells switched off.  ~3 minutes on 16 cores; the pytest subset is tests/test_me_host_emu_cpu.py."""
import sys, ctypes, subprocess, os, itertools, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import test_me_host_emu_cpu as t
ROOT='/root/repo'
libs={}
for name,flags in {"default":[], "no_vcell_reuse":["-DEMU_VCELL_REUSE_OFF=1"]}.items():
    so='/tmp/emu_sweep_%s.so'%name
    subprocess.run(["g++","-O1","-std=c++17","-shared","-fPIC","-pthread","-Wno-unknown-pragmas"]+flags+["-I",ROOT+"/tests/host_emu","-I",ROOT+"/x265-yuuki-asuna_b200/csrc","-I",ROOT+"/include","-o",so,ROOT+"/tests/host_emu/me_frame_emu.cpp"],check=True)
    libs[name]=ctypes.CDLL(so)
t0=time.time(); n=0
for name,L in libs.items():
    for depth in (8,10):
        for method in (0,1,2,3):
            for subme in range(8):
                merange = (16,24,57,32)[(method+subme)%4] if depth==8 else (16,24,40,32)[(method+subme)%4]
                t._check(L, depth, method, subme, merange, seed=5000+depth*100+method*10+subme, ctus=[(0,0),(1,0)])
                n+=1
    # FULL search, small range
    for subme in (0,2,5):
        t._check(L, 8, 5, subme, 8, seed=7000+subme, ctus=[(1,1)]); n+=1
    print(name, 'ok', n, 'cases', round(time.time()-t0,1),'s', flush=True)
