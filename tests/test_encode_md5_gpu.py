"""GPU: whole-encode drop-in proof (regression-tests.txt:3-7 convention: bit-exact bitstreams).
The reference's own CLI (source/x265.cpp + x265cli.cpp + abrEncApp.cpp, unmodified, linked by oracle/Makefile) encodes the same
synthetic clip twice: once with the C primitive table (oracle/_ref/x265_c_N) and once with common/primitives.cpp compiled as a
build with ENABLE_ASSEMBLY compiles it, so that x265_setup_primitives() (primitives.cpp:248-285, reached from
x265_encoder_open, api.cpp:76-260) installs OUR setupAssemblyPrimitives table (oracle/_ref/x265_b200_N = adapter +
libx265b200.so).  The two .hevc files must be identical: every installed slot then ran inside the real encoder -- WPP rows,
frame threads and lookahead workers calling concurrently, setupAliasPrimitives on top, the [ALIGNED] variants chosen by the
callers (quant.cpp:549, predict.cpp:289) -- and produced what the C table produces."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _blur(a, k=5):
    for ax in (0, 1):
        c = np.cumsum(a, axis=ax)
        pad = [(0, 0), (0, 0)]; pad[ax] = (k, 0)
        c = np.pad(c, pad)
        n = a.shape[ax]
        a = (np.take(c, np.arange(k, n + k), axis=ax) - np.take(c, np.arange(0, n), axis=ax)) / k
    return a


def write_clip(path, W, H, frames, depth, seed):
    """4:2:0 planar clip: textured luma under global motion + noise, gently varying chroma (8-bit bytes or 16-bit LE samples)"""
    rng = np.random.default_rng(seed)
    big = _blur(rng.uniform(0, 255, (H + 64, W + 64)))
    big = (big - big.min()) / (big.max() - big.min()) * 255
    cbig = _blur(rng.uniform(96, 160, (H // 2 + 32, W // 2 + 32)), 7)
    scale, pmax, dt = 1 << (depth - 8), (1 << depth) - 1, (np.uint16 if depth > 8 else np.uint8)
    x = y = 24
    with open(path, "wb") as f:
        for i in range(frames):
            x = int(np.clip(x + rng.integers(-5, 6), 0, 63)); y = int(np.clip(y + rng.integers(-5, 6), 0, 63))
            Y = np.clip(np.rint((big[y:y + H, x:x + W] + rng.normal(0, 2.0, (H, W))) * scale), 0, pmax).astype(dt)
            f.write(Y.tobytes())
            for k in range(2):
                C = np.clip(np.rint((cbig[y // 2 + 3 * k:y // 2 + 3 * k + H // 2, x // 2:x // 2 + W // 2] + rng.normal(0, 1.0, (H // 2, W // 2))) * scale), 0, pmax).astype(dt)
                f.write(C.tobytes())


def encode(exe, clip, W, H, frames, depth, out, extra):
    cmd = [os.path.join(REF, exe), "--input", clip, "--input-res", "%dx%d" % (W, H), "--fps", "30", "--input-depth", str(depth),
           "--frames", str(frames), "-o", out] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (cmd, r.stderr[-2000:])
    assert os.path.getsize(out) > 200, r.stderr[-500:]
    return hashlib.md5(open(out, "rb").read()).hexdigest(), r.stderr


@pytest.mark.parametrize("depth,W,H,frames,extra", [
    (8, 320, 180, 4, ["--preset", "medium", "--pools", "4", "-F", "2"]),                    # the case VERDICT r01 names
    (10, 192, 128, 3, ["--preset", "medium", "--pools", "4", "-F", "2"]),
    (8, 192, 128, 3, ["--preset", "slow", "--pools", "3", "-F", "1", "--me", "star"]),      # rect / AMP PUs, subme 3 chroma SATD, rdoq
    (8, 192, 128, 3, ["--preset", "ultrafast", "--pools", "2", "-F", "1"]),                 # CTU 32, DIA, no SAO
])
def test_bitstream_md5_equal_with_gpu_table(tmp_path, depth, W, H, frames, extra):
    for exe in ("x265_c_%d" % depth, "x265_b200_%d" % depth):
        assert os.path.exists(os.path.join(REF, exe)), "oracle/_ref/%s missing: run __graft_entry__.build() where /root/reference exists" % exe
    clip = str(tmp_path / "clip.yuv")
    write_clip(clip, W, H, frames, depth, seed=31 + depth)
    md5_c, _ = encode("x265_c_%d" % depth, clip, W, H, frames, depth, str(tmp_path / "c.hevc"), extra)
    md5_g, log = encode("x265_b200_%d" % depth, clip, W, H, frames, depth, str(tmp_path / "g.hevc"), extra)
    assert md5_c == md5_g, "bitstreams differ: C table %s, B200 table %s\n%s" % (md5_c, md5_g, log[-800:])
