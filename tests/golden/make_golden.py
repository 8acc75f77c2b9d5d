#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, read-only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference ships no known-answer tests for this path ("parity unpinned"), so these
files are the pin: every array below is an output of the reference's own code --

  * operators    : utils.mesh.{dist,area,cotangent_weights,dirac} + utils.graph.laplacian
                   (recipe of src/as_rigid_as_possible/add_laplacian.py:50-59)
  * batching     : utils.utils_pt.{sp_sparse_to_pt_sparse,sparse_diag_cat,sparse_cat}  (utils_pt.py:21-69)
  * SpMM         : torch.mm(sparse_coo, dense) exactly as at utils_pt.py:167,202,214
  * layers       : utils.utils_pt.{GraphConv1x1,LapResNet2,DirResNet2,DenseLapResNet2,AvgResNet2,
                   MlpResNet2} forward + backward (utils_pt.py:76-263)
  * models       : as_rigid_as_possible.models.{Model,DirModel} forward + loss gradient
                   (src/as_rigid_as_possible/models.py:21-52,108-152; loss main.py:225-226)

Nothing here is read at test time from /root/reference: the .npz files are committed.
Module parameters are filled by ``det_fill`` (tests/golden/det.py), a closed-form function of the
parameter name order, so fixtures do not depend on torch's RNG stream or init code.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF_SRC = os.environ.get("SN_REFERENCE_SRC", "/root/reference/src")
sys.dont_write_bytecode = True
sys.path.insert(0, REF_SRC)
sys.path.insert(0, os.path.join(REF_SRC, "as_rigid_as_possible"))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import warnings  # noqa: E402

warnings.filterwarnings("ignore")

import utils.graph as rgraph  # noqa: E402  (reference)
import utils.mesh as rmesh  # noqa: E402  (reference)
import utils.utils_pt as rutils  # noqa: E402  (reference)

from det import det_fill, det_tensor  # noqa: E402
from surfacenetworks_b200 import geometry  # noqa: E402  (mesh generator only)


def ref_operators(V, F):
    dist = rmesh.dist(V, F)
    areas = rmesh.area(F, dist)
    W, A = rmesh.cotangent_weights(F, areas, dist)
    L = rgraph.laplacian(W, symmetric=False, normalized=False)
    L = A * L
    Di, DiA = rmesh.dirac(V, F)
    return L.astype("float32"), Di.astype("float32"), DiA.astype("float32")


def coo_arrays(m, prefix):
    m = m.tocoo()
    return {prefix + "_row": m.row.astype(np.int64), prefix + "_col": m.col.astype(np.int64),
            prefix + "_val": m.data.astype(np.float32), prefix + "_shape": np.array(m.shape, dtype=np.int64)}


def pt_coo_arrays(t, prefix):
    return {prefix + "_idx": t._indices().numpy().copy(), prefix + "_val": t._values().numpy().copy(),
            prefix + "_shape": np.array(t.size(), dtype=np.int64)}


def grads_of(module):
    return {k: p.grad.detach().numpy().copy() for k, p in module.named_parameters()}


def run_block(module, args, tensor_inputs, out_weights):
    """Forward + backward of a reference module; loss = sum_i <out_i, w_i> so grads are non-trivial."""
    module.train()
    for t in tensor_inputs:
        t.requires_grad_(True)
    outs = module(*args)
    if not isinstance(outs, tuple):
        outs = (outs,)
    loss = sum((o * w).sum() for o, w in zip(outs, out_weights))
    loss.backward()
    rec = {"out%d" % i: o.detach().numpy().copy() for i, o in enumerate(outs)}
    for i, t in enumerate(tensor_inputs):
        rec["gin%d" % i] = t.grad.detach().numpy().copy()
    for k, g in grads_of(module).items():
        rec["gparam." + k] = g
    for k, b in module.named_buffers():
        rec["buf." + k] = b.detach().numpy().copy()
    return rec


def save(name, **arrays):
    only = os.environ.get("SN_GOLDEN_ONLY")          # e.g. SN_GOLDEN_ONLY=siamese.npz: leave the other files untouched
    if only and name not in only.split(","):
        return
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote %-22s %7.1f kB  (%d arrays)" % (name, os.path.getsize(path) / 1e3, len(arrays)))


def main():
    torch.set_num_threads(1)
    # ------------------------------------------------------------------ operators
    Vc, Fc = geometry.read_ply_ascii(os.path.join(os.path.dirname(REF_SRC), "meshes", "cube.ply"))
    meshes = {"cube": (Vc, Fc), "s60": geometry.synth_mesh(60, 0), "s45": geometry.synth_mesh(45, 1)}
    ops = {}
    arrays = {}
    for name, (V, F) in meshes.items():
        L, Di, DiA = ref_operators(V, F)
        ops[name] = (L, Di, DiA)
        arrays[name + "_V"] = V
        arrays[name + "_F"] = F
        arrays.update(coo_arrays(L, name + "_L"))
        arrays.update(coo_arrays(Di, name + "_Di"))
        arrays.update(coo_arrays(DiA, name + "_DiA"))
    save("operators.npz", **arrays)

    # ------------------------------------------------------------------ batching helpers (ragged batch of 2)
    (L0, Di0, DiA0), (L1, Di1, DiA1) = ops["s60"], ops["s45"]
    nv = max(L0.shape[0], L1.shape[0])
    nf = max(Di0.shape[0], Di1.shape[0]) // 4
    ptL = [rutils.sp_sparse_to_pt_sparse(m) for m in (L0, L1)]
    ptDi = [rutils.sp_sparse_to_pt_sparse(m) for m in (Di0, Di1)]
    ptDiA = [rutils.sp_sparse_to_pt_sparse(m) for m in (DiA0, DiA1)]
    Lb = rutils.sparse_diag_cat(ptL, nv, nv)
    Dib = rutils.sparse_diag_cat(ptDi, 4 * nf, 4 * nv)
    DiAb = rutils.sparse_diag_cat(ptDiA, 4 * nv, 4 * nf)
    L3 = rutils.sparse_cat(ptL, nv, nv)
    Di3 = rutils.sparse_cat(ptDi, 4 * nf, 4 * nv)
    arrays = {"nv": np.int64(nv), "nf": np.int64(nf)}
    arrays.update(pt_coo_arrays(ptL[1], "pt_L1"))
    arrays.update(pt_coo_arrays(Lb, "diag_L"))
    arrays.update(pt_coo_arrays(Dib, "diag_Di"))
    arrays.update(pt_coo_arrays(DiAb, "diag_DiA"))
    arrays.update(pt_coo_arrays(L3, "cat_L"))
    arrays.update(pt_coo_arrays(Di3, "cat_Di"))
    save("batching.npz", **arrays)

    # ------------------------------------------------------------------ SpMM known answers
    arrays = {}
    Lc, Dic, DiAc = [rutils.sp_sparse_to_pt_sparse(m).coalesce() for m in ops["cube"]]
    x = det_tensor((8, 16), 11)
    f = det_tensor((12, 16), 12)
    arrays["cube_x"] = x.numpy()
    arrays["cube_f"] = f.numpy()
    arrays["cube_Lx"] = torch.mm(Lc, x).numpy()
    arrays["cube_Dix"] = torch.mm(Dic, x.view(8 * 4, 4)).view(12, 16).numpy()       # utils_pt.py:201-203
    arrays["cube_DiAf"] = torch.mm(DiAc, f.view(12 * 4, 4)).view(8, 16).numpy()     # utils_pt.py:213-215
    for C in (4, 32, 40, 128):
        xb = det_tensor((2 * nv, C), 20 + C)
        fb = det_tensor((2 * nf, C), 21 + C)
        arrays["b_x%d" % C] = xb.numpy()
        arrays["b_f%d" % C] = fb.numpy()
        arrays["b_Lx%d" % C] = torch.mm(Lb, xb).numpy()
        arrays["b_LTx%d" % C] = torch.mm(Lb.t().coalesce(), xb).numpy()
        arrays["b_Dix%d" % C] = torch.mm(Dib, xb.view(2 * nv * 4, C // 4)).view(2 * nf, C).numpy()
        arrays["b_DiAf%d" % C] = torch.mm(DiAb, fb.view(2 * nf * 4, C // 4)).view(2 * nv, C).numpy()
        arrays["b_DiTf%d" % C] = torch.mm(Dib.t().coalesce(), fb.view(2 * nf * 4, C // 4)).view(2 * nv, C).numpy()
    save("spmm.npz", **arrays)

    # ------------------------------------------------------------------ layer library, cube (cfg1) and ragged batch
    arrays = {}

    def record(tag, rec):
        for k, v in rec.items():
            arrays[tag + "/" + k] = v

    C = 16
    xin = det_tensor((1, 8, C), 31)
    fin = det_tensor((1, 12, C), 32)
    mask1 = torch.ones(1, 8, 1)
    m = det_fill(rutils.LapResNet2(C), 1)
    record("cube_lap", run_block(m, (Lc, mask1, xin), [xin], [det_tensor((1, 8, C), 33)]))
    xin = det_tensor((1, 8, C), 31)
    m = det_fill(rutils.DirResNet2(C), 2)
    record("cube_dir", run_block(m, (Dic, DiAc, xin, fin), [xin, fin],
                                 [det_tensor((1, 8, C), 34), det_tensor((1, 12, C), 35)]))

    C = 32
    B = 2
    mask = torch.zeros(B, nv, 1)
    mask[0, :L0.shape[0]] = 1
    mask[1, :L1.shape[0]] = 1
    arrays["mask"] = mask.numpy()
    wv = det_tensor((B, nv, C), 43)
    wf = det_tensor((B, nf, C), 44)

    def fresh():
        return det_tensor((B, nv, C), 41), det_tensor((B, nf, C), 42)

    xin, fin = fresh()
    record("b_lap", run_block(det_fill(rutils.LapResNet2(C), 3), (Lb, mask, xin), [xin], [wv]))
    xin, fin = fresh()
    record("b_dir", run_block(det_fill(rutils.DirResNet2(C), 4), (Dib, DiAb, xin, fin), [xin, fin], [wv, wf]))
    xin, fin = fresh()
    record("b_avg", run_block(det_fill(rutils.AvgResNet2(C), 5), (None, mask, xin), [xin], [wv]))
    xin, fin = fresh()
    record("b_mlp", run_block(det_fill(rutils.MlpResNet2(C), 6), (None, mask, xin), [xin], [wv]))
    xin, fin = fresh()
    Ldense = torch.stack([Lb.to_dense()[i * nv:(i + 1) * nv, i * nv:(i + 1) * nv] for i in range(B)])
    record("b_denselap", run_block(det_fill(rutils.DenseLapResNet2(C), 3), (Ldense, mask, xin), [xin], [wv]))
    xin, fin = fresh()
    record("b_lap_densearg", run_block(det_fill(rutils.LapResNet2(C), 3), (Ldense, mask, xin), [xin], [wv]))
    for bn in (None, "pre", "post"):
        xin, fin = fresh()
        record("b_conv_%s" % bn, run_block(det_fill(rutils.GraphConv1x1(C, 24, batch_norm=bn), 7), (xin,), [xin],
                                           [det_tensor((B, nv, 24), 45)]))
    xin, fin = fresh()
    record("b_gbn", run_block(det_fill(rutils.GraphBatchNorm(C), 8), (xin,), [xin], [wv]))
    arrays["b_global_average"] = rutils.global_average(det_tensor((B, nv, C), 41), mask).numpy()
    # eval-mode (running statistics) forward of the two hot blocks
    with torch.no_grad():
        m = det_fill(rutils.LapResNet2(C), 3).eval()
        arrays["b_lap_eval/out0"] = m(Lb, mask, det_tensor((B, nv, C), 41)).numpy()
        m = det_fill(rutils.DirResNet2(C), 4).eval()
        o = m(Dib, DiAb, det_tensor((B, nv, C), 41), det_tensor((B, nf, C), 42))
        arrays["b_dir_eval/out0"] = o[0].numpy()
        arrays["b_dir_eval/out1"] = o[1].numpy()
    save("layers.npz", **arrays)

    # ------------------------------------------------------------------ ARAP models (callers)
    import models as arap_models  # reference src/as_rigid_as_possible/models.py

    arrays = {}
    inputs = det_tensor((B, nv, 6), 51) * mask
    targets = det_tensor((B, nv, 120), 52) * mask
    arrays["inputs"] = inputs.numpy()
    arrays["targets"] = targets.numpy()
    for tag, make, seed, args in (("dir", arap_models.DirModel, 9, (Dib, DiAb, mask, inputs)),
                                  ("lap", lambda: arap_models.Model(15), 10, (Lb, mask, inputs))):
        model = det_fill(make(), seed, gain=0.25)
        model.train()
        out = model(*args)
        outm = out * mask.expand_as(out)
        loss = torch.nn.functional.smooth_l1_loss(outm, targets, reduction="sum") / B   # main.py:225-226
        loss.backward()
        arrays[tag + "/out"] = out.detach().numpy()
        arrays[tag + "/loss"] = loss.detach().numpy()
        # keep the fixture small: a few representative parameter gradients
        named = dict(model.named_parameters())
        for k in ("conv1.fc.weight", "rn0.bn_fc0.fc.weight", "rn0.bn_fc0.bn.weight", "rn0.bn_fc1.fc.bias",
                  "rn7.bn_fc0.fc.bias", "rn14.bn_fc1.bn.bias", "conv2.fc.weight"):
            arrays[tag + "/g." + k] = named[k].grad.numpy().copy()
        # the same reference modules in double precision: how far the reference's own fp32 result is from exact
        model64 = det_fill(make(), seed, gain=0.25).double().train()
        args64 = tuple(a.double() if torch.is_tensor(a) else a for a in args)
        out64 = model64(*args64)
        loss64 = torch.nn.functional.smooth_l1_loss(out64 * mask.double().expand_as(out64), targets.double(),
                                                    reduction="sum") / B
        loss64.backward()
        arrays[tag + "/out64"] = out64.detach().numpy()
        arrays[tag + "/loss64"] = loss64.detach().numpy()
        named64 = dict(model64.named_parameters())
        for k in [k for k in list(arrays) if k.startswith(tag + "/g.")]:
            arrays[k.replace("/g.", "/g64.")] = named64[k[len(tag) + 3:]].grad.numpy().copy()
    save("arap_models.npz", **arrays)

    # ------------------------------------------------------------------ other callers
    import importlib.util

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    np_models = load("ref_normal_models", os.path.join(REF_SRC, "normal_predict", "models.py"))
    vae_models = load("ref_vae_models", os.path.join(REF_SRC, "mesh_mnist", "models_vae.py"))
    arrays = {}
    x3 = det_tensor((B, nv, 3), 61) * mask
    arrays["x3"] = x3.numpy()
    xin32 = det_tensor((B, nv, 32), 62)

    def both_precisions(make, seed, args, tag, gain=0.25):
        for dt, suffix in ((torch.float32, ""), (torch.float64, "64")):
            m = det_fill(make(), seed, gain=gain).to(dt).train()
            a = tuple((t.to(dt) if torch.is_tensor(t) and t.is_floating_point() else t) for t in args)
            a = tuple(tuple(u.to(dt) for u in t) if isinstance(t, tuple) else t for t in a)
            out = m(*a)
            outs = out if isinstance(out, tuple) else (out,)
            for i, o in enumerate(outs):
                arrays["%s/out%d%s" % (tag, i, suffix)] = o.detach().numpy()

    # normal_predict _LapResNet2 (models.py:447-477): generalised Laplacian block
    both_precisions(lambda: np_models._LapResNet2(32, 64, inner_layers=3), 11, (Lb, mask, xin32), "lapgen_32_64_3", gain=0.5)
    both_precisions(lambda: np_models._LapResNet2(32), 12, (Lb, mask, xin32), "lapgen_32", gain=0.5)
    both_precisions(lambda: np_models._LapResNet2(32, 16, inner_layers=1), 15, (Lb, mask, xin32), "lapgen_32_16_1", gain=0.5)
    # normal_predict DirDeepModel (models.py:234-274), 4 blocks
    both_precisions(lambda: np_models.DirDeepModel(3, 1, layers=4), 13, ((Dib, DiAb), mask, x3), "dirdeep4")
    # mesh_mnist VAE LapEncoder (models_vae.py:22-51), 5 x LapResNet2(128)
    both_precisions(lambda: vae_models.LapEncoder(), 14, (x3, Lb, mask), "lapencoder")
    save("callers.npz", **arrays)

    # ------------------------------------------------------------------ dense_correspondence SiameseModel (8(f) f4)
    # models.py:184-203: shared tower on both shapes + torch.bmm(FA, FB^T).  Only the 'lap' tower runs in the
    # reference (the Dirac tower reads DiA.size(2) on a 2-D operator).
    dc_models = load("ref_dc_models", os.path.join(REF_SRC, "dense_correspondence", "models.py"))
    arrays = {}
    xb3 = det_tensor((B, nv, 3), 63) * mask
    arrays["xa"], arrays["xb"] = x3.numpy(), xb3.numpy()
    both_precisions(lambda: dc_models.SiameseModel("lap", 3), 16, ((Lb, mask), (Lb, mask), x3, xb3), "siamese_lap3")
    both_precisions(lambda: dc_models.Model(5), 17, (Lb, mask, x3), "dclap5")
    save("siamese.npz", **arrays)

    # ------------------------------------------------------------------ backward of the same callers (round 2)
    # loss = sum_i <out_i, w_i> with closed-form weights; recorded per model and precision: outputs, the gradient of the
    # feature input and every parameter gradient (fp64 runs stored as float32: they only set the noise scale of
    # tests/test_gpu_layers.py::within_reference_noise).
    arrays = {}

    def with_grads(make, seed, args, grad_arg, tag, gain=0.25, keep=None):
        """fp32 run: arrays; fp64 run: per-tensor (max|g32 - g64|, max|g64|) pairs -- the reference's own rounding noise, the
        unit of the model-level criterion.  ``keep``: parameter-name prefixes whose gradients are stored (None = all)."""
        rec = {}
        for dt in (torch.float32, torch.float64):
            m = det_fill(make(), seed, gain=gain).to(dt).train()

            def cast(t):
                if isinstance(t, tuple):
                    return tuple(cast(u) for u in t)
                return t.to(dt) if torch.is_tensor(t) and t.is_floating_point() else t
            a = [cast(t) for t in args]
            a[grad_arg] = a[grad_arg].clone().requires_grad_(True)
            out = m(*a)
            outs = out if isinstance(out, tuple) else (out,)
            loss = sum((o * det_tensor(tuple(o.shape), 900 + 7 * i + seed).to(dt)).sum() for i, o in enumerate(outs))
            loss.backward()
            cur = {"out%d" % i: o.detach().numpy() for i, o in enumerate(outs)}
            cur["gin"] = a[grad_arg].grad.numpy()
            for k, p_ in m.named_parameters():
                if p_.grad is not None and (keep is None or k.startswith(tuple(keep))):
                    cur["g." + k] = p_.grad.numpy()
            rec[dt] = cur
        for k, v32 in rec[torch.float32].items():
            v64 = rec[torch.float64][k]
            arrays["%s/%s" % (tag, k)] = v32.astype(np.float32)
            arrays["%s/%s#noise" % (tag, k)] = np.array([np.abs(v32.astype(np.float64) - v64).max(), np.abs(v64).max()])

    with_grads(lambda: np_models._LapResNet2(32, 64, inner_layers=3), 11, (Lb, mask, xin32), 2, "lapgen_32_64_3", gain=0.5)
    with_grads(lambda: np_models._LapResNet2(32), 12, (Lb, mask, xin32), 2, "lapgen_32", gain=0.5)
    with_grads(lambda: np_models._LapResNet2(32, 16, inner_layers=1), 15, (Lb, mask, xin32), 2, "lapgen_32_16_1", gain=0.5)
    big = ("conv1.", "rn0.", "conv2.", "bn_conv2.", "fc_mu.", "fc_logvar.", "model.conv1.", "model.rn0.", "model.conv2.")
    with_grads(lambda: np_models.DirDeepModel(3, 1, layers=4), 13, ((Dib, DiAb), mask, x3), 2, "dirdeep4", keep=big + ("rn3.",))
    with_grads(lambda: vae_models.LapEncoder(), 14, (x3, Lb, mask), 0, "lapencoder", keep=big + ("rn4.",))
    with_grads(lambda: dc_models.SiameseModel("lap", 3), 16, ((Lb, mask), (Lb, mask), x3, xb3), 2, "siamese_lap3",
               keep=big + ("model.rn2.",))
    save("callers_grads.npz", **arrays)


if __name__ == "__main__":
    main()
