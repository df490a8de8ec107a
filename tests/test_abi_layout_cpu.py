"""CPU: the Python mirrors of the C-ABI records (numpy dtypes / ctypes Structures in capi.py) must have exactly the layout the C
compiler gives the structs of include/x265b200.h -- sizes and field offsets are printed by a probe compiled with gcc."""
import ctypes
import importlib
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("x265-yuuki-asuna_b200")

PROBE = r'''
#include <stdio.h>
#include <stddef.h>
#include "x265b200.h"
#define S(T) printf(#T " size %zu\n", sizeof(T))
#define F(T, f) printf(#T " " #f " %zu\n", offsetof(T, f))
int main(void)
{
    S(x265b200_me_job); F(x265b200_me_job, puX); F(x265b200_me_job, w); F(x265b200_me_job, mvminX); F(x265b200_me_job, mvpX); F(x265b200_me_job, numCand);
    F(x265b200_me_job, mvc); F(x265b200_me_job, refIdx); F(x265b200_me_job, outMvX); F(x265b200_me_job, outCost);
    S(x265b200_glue_job); S(x265b200_interp_job); F(x265b200_interp_job, idxX); S(x265b200_intra_job); F(x265b200_intra_job, mode);
    S(x265b200_la_triple); F(x265b200_la_triple, doSearch); F(x265b200_la_triple, mvSlot); F(x265b200_la_triple, weightIdx0); F(x265b200_la_triple, weightPlanes0);
    S(x265b200_la_weight_job); F(x265b200_la_weight_job, intraCost); F(x265b200_la_weight_job, refBuffer); F(x265b200_la_weight_job, weighted); F(x265b200_la_weight_job, fencSum); F(x265b200_la_weight_job, refSsd);
    S(x265b200_la_weight); F(x265b200_la_weight, inputOffset); F(x265b200_la_weight, origscore); F(x265b200_la_weight, score);
    S(x265b200_ads_job); F(x265b200_ads_job, thresh); F(x265b200_ads_job, encDC);
    S(x265b200_mc_job); F(x265b200_mc_job, cuX); F(x265b200_mc_job, refIdx); F(x265b200_mc_job, mv);
    S(x265b200_mc_weight); S(x265b200_mc_desc); F(x265b200_mc_desc, refs); F(x265b200_mc_desc, predY); F(x265b200_mc_desc, predStrideY); F(x265b200_mc_desc, weights);
    S(x265b200_sao_job); F(x265b200_sao_job, buf0); F(x265b200_sao_job, offsetOff); F(x265b200_sao_job, width); F(x265b200_sao_job, startX);
    S(x265b200_deblock_job); F(x265b200_deblock_job, offset); F(x265b200_deblock_job, tcP); F(x265b200_deblock_job, maskQ);
    S(x265b200_me_chroma); F(x265b200_me_chroma, fencCb); F(x265b200_me_chroma, fencStrideC); F(x265b200_me_chroma, refCbPlanes); F(x265b200_me_chroma, refStrideC);
    S(x265b200_la_hme); F(x265b200_la_hme, lowerStride); F(x265b200_la_hme, width4); F(x265b200_la_hme, lowerMvPool); F(x265b200_la_hme, searchMethod); F(x265b200_la_hme, range);
    S(x265b200_sad_group); F(x265b200_sad_group, ref);
    S(x265b200_interp_seg); F(x265b200_interp_seg, isRowExt); F(x265b200_interp_seg, src); F(x265b200_interp_seg, dstStride); F(x265b200_interp_seg, jobs); F(x265b200_interp_seg, n);
    S(x265b200_me_frame_params); F(x265b200_me_frame_params, minCuSize); F(x265b200_me_frame_params, picWidth); F(x265b200_me_frame_params, chromaMarginX); F(x265b200_me_frame_params, numRefs);
    F(x265b200_me_frame_params, merange); F(x265b200_me_frame_params, maxCand); F(x265b200_me_frame_params, sliceTotalRows); F(x265b200_me_frame_params, refLagPixels);
    S(x265b200_me_frame_planes); F(x265b200_me_frame_planes, curCr); F(x265b200_me_frame_planes, curStrideC); F(x265b200_me_frame_planes, refY); F(x265b200_me_frame_planes, refStrideC);
    return 0;
}
'''


def _probe():
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "probe.c"), os.path.join(d, "probe")
        open(src, "w").write(PROBE)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    res = {}
    for line in out.splitlines():
        parts = line.split()
        res[(parts[0], parts[1])] = int(parts[2])
    return res


def test_record_layouts_match_the_header():
    c = _probe()
    dtypes = {"x265b200_me_job": pkg.ME_JOB, "x265b200_interp_job": pkg.INTERP_JOB, "x265b200_intra_job": pkg.INTRA_JOB,
              "x265b200_la_triple": pkg.LA_TRIPLE, "x265b200_la_weight_job": pkg.LA_WEIGHT_JOB, "x265b200_la_weight": pkg.LA_WEIGHT, "x265b200_ads_job": pkg.ADS_JOB, "x265b200_mc_job": pkg.MC_JOB,
              "x265b200_mc_weight": pkg.MC_WEIGHT, "x265b200_sad_group": pkg.SAD_GROUP, "x265b200_sao_job": pkg.SAO_JOB, "x265b200_deblock_job": pkg.DEBLOCK_JOB}
    if hasattr(pkg, "GLUE_JOB"):
        dtypes["x265b200_glue_job"] = pkg.GLUE_JOB
    for name, dt in dtypes.items():
        dt = np.dtype(dt)
        assert dt.itemsize == c[(name, "size")], (name, dt.itemsize, c[(name, "size")])
        for (n, f), off in c.items():
            if n == name and f != "size":
                assert dt.fields[f][1] == off, (name, f, dt.fields[f][1], off)
    structs = {"x265b200_mc_desc": pkg.MC_DESC, "x265b200_me_chroma": pkg.ME_CHROMA, "x265b200_la_hme": pkg.LA_HME, "x265b200_interp_seg": pkg.INTERP_SEG,
               "x265b200_me_frame_params": pkg.ME_FRAME_PARAMS, "x265b200_me_frame_planes": pkg.ME_FRAME_PLANES}
    for name, st in structs.items():
        assert ctypes.sizeof(st) == c[(name, "size")], (name, ctypes.sizeof(st), c[(name, "size")])
        for (n, f), off in c.items():
            if n == name and f != "size":
                assert getattr(st, f).offset == off, (name, f, getattr(st, f).offset, off)
    assert pkg.ME_FRAME_PARAMS.lambda_.offset == _probe_lambda_offset()


def _probe_lambda_offset():
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "p.c"), os.path.join(d, "p")
        open(src, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "x265b200.h"\nint main(void){printf("%zu", offsetof(x265b200_me_frame_params, lambda));return 0;}')
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src], check=True)
        return int(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)
