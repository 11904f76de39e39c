"""The GraspBalance point-op pipeline: the exact sequence of hot-path operator calls one forward + backward of the
reference model issues per batch of 20k-point scenes, with random tensors standing in for the
dense MLP outputs between them (the MLPs/heads are cuDNN/cuBLAS work and out of scope).

Stages and the reference call sites they reproduce:
  SA1..SA4   PointnetSAModuleVotes.forward (PointNet/pointnet2_modules.py:148-188): furthest_point_sample ->
             gather_operation -> QueryAndGroup (ball_query + grouping_operation on xyz and features), variant A;
             npoint/radius/nsample from TrainModel/drp.py:161-246.
  InvResMLP  3+6+3+3 LocalAggregation blocks (drp.py:62-67): group.QueryAndGroup (ball_query + two grouping_operations),
             variant B; radii/nsample from drp.py:169-247.
  FP1, FP2   PointnetFPModule.forward (pointnet2_modules.py:407-435): three_nn -> weights -> three_interpolate.
  UP         seed features back to the full cloud (TrainModel/graspbalance.py:37-41): three_nn + three_interpolate, C=256.
  CROP       4 GraspWidthGrouping x 4 depths = 16 CylinderQueryAndGroup calls in the reference (TrainModel/modules.py:104-124,
             graspbalance.py:84-87,123: radii 0.08 x {.25,.5,.75,1}, hmin -0.02, hmax {.01,.02,.03,.04}, nsample 64); here
             each GraspWidthGrouping is one multi-depth scan + one grouped-coordinate launch (graspbalance_b200/modules.py,
             bit-identical to the loop; `fused_crops=False` runs the reference's 16 separate calls).
  COLLISION  ModelFreeCollisionDetector.detect's occupancy test for 1024 grasps per scene (collision_detector.py:16-64).
Every grouped / interpolated feature tensor is back-propagated with a random upstream gradient through the reference's
autograd Functions (GroupingOperation / GatherOperation / ThreeInterpolate backward).

Everything here goes through the public drop-in API (pointnet2_utils, group, upsampling, collision_detector).
"""
import math

import numpy as np
import os

import torch

from . import group as gb_group
from . import pointnet2_utils as pu
from .collision_detector import collision_counts, collision_counts_batched
from .modules import GraspWidthGrouping, multi_scale_group

# (npoint, radius, nsample, C_in) of the four SA modules and (blocks, C, radius, nsample) of the InvResMLP groups
SA_SPECS = [(2048, 0.04, 64, 0), (1024, 0.1, 32, 128), (512, 0.2, 16, 256), (256, 0.3, 16, 256)]
IRM_SPECS = [(3, 128, 0.08, 64), (6, 256, 0.2, 32), (3, 256, 0.4, 16), (3, 256, 0.6, 16)]
CROP_RADII = [0.08 * s for s in (0.25, 0.5, 0.75, 1.0)]
CROP_HMAX = [0.01, 0.02, 0.03, 0.04]
CROP_HMIN = -0.02
NUM_SEED = 1024
NUM_GRASP = 1024


def _randn(shape, gen, device):
    return torch.randn(shape, generator=gen, device=device, dtype=torch.float32)


class OpPipeline:
    """Holds the stand-in feature / gradient tensors (allocated once, outside any timed region) and runs the chain."""
    # FP modules as one launch (gb_three_interpolation).  Off: measured 1.24 ms per step against 0.83 ms for three_nn_weights +
    # three_interpolate (B200, 32 scenes) -- the brute-force search inside a 2-CTA-per-SM kernel loses more than the saved
    # 36 bytes per point of idx / weight traffic; it stays the path that writes nothing but the output (inference).
    fused_fp = False
    # Schedule of the groupers' backward launches: right after each forward (False) or after every forward of the step, last
    # grouping first -- the order loss.backward() visits them in (True).
    # CTAs per scene of the background sampling chain (gb_fps_xyz_hint; 0 = the latency-optimal shape)
    background_cluster = int(os.environ.get("GB_FPS_BG_CLUSTER", "2"))
    backward_last = os.environ.get("GB_BACKWARD_LAST", "1") == "1"  # B200, 32 scenes: 10.62 ms per step against 10.67

    def __init__(self, batch, n_points=20000, device="cuda", seed=0, backward=True, overlap=True, fused_crops=True, fused_sampling=True, batched_collision=True):
        self.B, self.N, self.device, self.backward = batch, n_points, torch.device(device), backward
        # overlap: the sampling chain (4 x FPS + gather: latency-bound, a few warps per SM, depends on xyz only) and the
        # collision tests (independent of everything else) run on side streams next to the bandwidth-bound grouping work
        self.overlap = overlap and self.device.type == "cuda"
        self._main_stream = None
        if self.overlap:
            prio = lambda name: int(os.environ.get(name, "0"))  # CUDA stream priority (0 = default, -1.. = higher)
            main_prio = int(os.environ.get("GB_PRIO_MAIN", "-1"))
            self._main_stream = torch.cuda.Stream(self.device, priority=main_prio) if main_prio else None
            self._fps_stream = torch.cuda.Stream(self.device, priority=prio("GB_PRIO_FPS"))
            self._col_stream = torch.cuda.Stream(self.device, priority=prio("GB_PRIO_COL"))
            self._aux_stream = torch.cuda.Stream(self.device, priority=prio("GB_PRIO_AUX"))  # grasp crops + FP/up-sampling chain (compute-bound scans)
        gen = torch.Generator(device=self.device).manual_seed(seed)
        B = batch
        self.sa_groupers = [pu.QueryAndGroup(r, ns, use_xyz=True, ret_grouped_xyz=True, normalize_xyz=True)
                            for (_, r, ns, _) in SA_SPECS]
        self.irm_groupers = [gb_group.QueryAndGroup(r, ns) for (_, _, r, ns) in IRM_SPECS]
        self.fused_crops = fused_crops
        self.fused_sampling = fused_sampling
        self.batched_collision = batched_collision
        self.crop_modules = [GraspWidthGrouping(64, 3, cylinder_radius=r, hmin=CROP_HMIN, hmax_list=CROP_HMAX, mlps=torch.nn.Identity())
                             for r in CROP_RADII]
        # stand-ins for MLP outputs (features entering each stage) and for upstream gradients
        self.sa_in_feats = [None] + [_randn((B, c, SA_SPECS[i - 1][0]), gen, self.device) for i, (_, _, _, c) in
                                     enumerate(SA_SPECS) if i > 0]
        self.irm_feats = [_randn((B, c, SA_SPECS[i][0]), gen, self.device) for i, (_, c, _, _) in enumerate(IRM_SPECS)]
        # upstream gradient of the whole (3+C)-channel grouped tensor, as the SharedMLP's backward hands it over
        self.sa_grads = [None] + [_randn((B, 3 + SA_SPECS[i][3], SA_SPECS[i][0], SA_SPECS[i][2]), gen, self.device)
                                  for i in range(1, 4)]
        self.irm_grads = [_randn((B, c, SA_SPECS[i][0], ns), gen, self.device) for i, (_, c, _, ns) in enumerate(IRM_SPECS)]
        self.fp_feats = [_randn((B, 256, 256), gen, self.device), _randn((B, 256, 512), gen, self.device),
                         _randn((B, 256, 1024), gen, self.device)]
        self.fp_grads = [_randn((B, 256, 512), gen, self.device), _randn((B, 256, 1024), gen, self.device),
                         _randn((B, 256, n_points), gen, self.device)]
        for t in self.sa_in_feats[1:] + self.irm_feats + self.fp_feats:
            t.requires_grad_(backward)

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _interp(unknown, known, feats, grad, collect=None, tag=""):
        # three_nn + weights of pointnet2_modules.py:413-416 (sqrt, +1e-8, reciprocal, sum, divide) + three_interpolate: one launch
        if collect is None and OpPipeline.fused_fp:
            out = pu.three_interpolation(unknown, known, feats)
        else:
            _, idx, weight = pu.three_nn_weights(unknown, known)
            out = pu.three_interpolate(feats, idx, weight)
        if grad is not None:
            out.backward(grad)
        if collect is not None:
            collect[tag + "_idx"], collect[tag + "_weight"], collect[tag + "_out"] = idx, weight, out.detach()
        return out

    def _crops(self, xyz, view_rot, sa2_xyz, out, collect=None):
        # ---- grasp crop: 4 radii x 4 depths cylinder query + group (seeds = fp2_xyz: 1024 points, drp.py:301-303) ----
        parts = []  # small slices: the checksum only gives the step a result to return
        if self.fused_crops:
            # WidthGroup1..4 of graspbalance.py:104-107 differ in the radius only: one scan for all 4 x 4 cylinders
            groups = multi_scale_group(self.crop_modules, sa2_xyz, xyz, view_rot)
            parts = [g[:, :, :16] for g in groups]  # seeds 0..3 x 4 depths
            if collect is not None:  # [B,3,seed*D+d,ns] -> one [B,3,seed,ns] per (radius k, depth d)
                for k, g in enumerate(groups):
                    g5 = g.view(g.shape[0], 3, NUM_SEED, len(CROP_HMAX), g.shape[-1])
                    for d in range(len(CROP_HMAX)):
                        collect[f"crop{k}_{d}_xyz"] = g5[:, :, :, d, :]
        else:
            for k, mod in enumerate(self.crop_modules):
                full = [gq(xyz, sa2_xyz, view_rot) for gq in mod.groupers]  # 4 x [B,3,1024,64]
                parts += [g[:, :, :4] for g in full]
                if collect is not None:
                    for d, g in enumerate(full):
                        collect[f"crop{k}_{d}_xyz"] = g
        out["crop_checksum"] = torch.cat([p.reshape(-1) for p in parts]).sum()

    def _interpolation(self, xyz, sa2_xyz, sa3_xyz, sa4_xyz, out, collect=None):
        bw = self.backward
        # ---- FP modules + up-sampling of the seed features to the full cloud ----
        self._interp(sa3_xyz, sa4_xyz, self.fp_feats[0], self.fp_grads[0] if bw else None, collect, "fp0")
        self._interp(sa2_xyz, sa3_xyz, self.fp_feats[1], self.fp_grads[1] if bw else None, collect, "fp1")
        up = self._interp(xyz, sa2_xyz, self.fp_feats[2], self.fp_grads[2] if bw else None, collect, "fp2")
        out["up_checksum"] = up[:, :4, :256].sum()

    def alloc_samples(self):
        """Static buffers for one sampling chain: [(inds [B,npoint] i32, new_xyz [B,npoint,3] f32)] per SA level."""
        return [(torch.empty((self.B, m), dtype=torch.int32, device=self.device),
                 torch.empty((self.B, m, 3), dtype=torch.float32, device=self.device)) for (m, _, _, _) in SA_SPECS]

    def sampling_chain(self, xyz, into, background=False):
        """The four furthest_point_sample + gather_operation calls of a step (pointnet2_modules.py:151-158), which depend on
        the coordinates only, on the CURRENT stream, results written into the buffers of alloc_samples().  background: the
        chain runs beside a whole step of other kernels, so it is launched with at most two CTAs per scene -- slower rounds
        on 64 SMs instead of faster ones on 128, and the register-heavy kernels of the step keep the other SMs to themselves."""
        cur = xyz
        for (inds_buf, xyz_buf), (npoint, _, _, _) in zip(into, SA_SPECS):
            # (a shard of fewer than 8 scenes has no step long enough to hide slower rounds: automatic shape)
            inds, new_xyz = pu.furthest_point_sample_xyz(cur, npoint, self.background_cluster if background and self.B >= 8 else 0)
            inds_buf.copy_(inds)
            xyz_buf.copy_(new_xyz)
            cur = xyz_buf

    def run(self, xyz, view_rot, grasps=None, collect=None, samples=None, prefetch=None):
        """One step (see _run).  With the overlapped schedule the step's main line -- the bandwidth-bound group forward /
        backward launches, its critical path -- is issued on a stream of HIGHER PRIORITY than the side streams (sampling
        chain, collision tests, crops + interpolation), forked from and joined back into the caller's current stream: when
        SMs free up the block scheduler places the main line's CTAs first and the compute-bound scans fill what is left
        (B200, 32 scenes: 10.36 -> 10.17 ms per step; the side streams above the main line instead: 10.53)."""
        if not (self.overlap and self._main_stream is not None):
            return self._run(xyz, view_rot, grasps, collect, samples, prefetch)
        caller = torch.cuda.current_stream(self.device)
        self._main_stream.wait_stream(caller)
        with torch.cuda.stream(self._main_stream):
            out = self._run(xyz, view_rot, grasps, collect, samples, prefetch)
        caller.wait_stream(self._main_stream)
        for t in out.values():  # allocated in the main line's pool, consumed on the caller's stream
            t.record_stream(caller)
        return out

    def _run(self, xyz, view_rot, grasps=None, collect=None, samples=None, prefetch=None):
        """xyz [B,N,3] f32 CUDA; view_rot [B,1024,3,3] f32 CUDA (approach frames of the seeds); grasps = optional dict of
        per-scene fp64 CUDA tensors {scene_points: list of [N'_b,3], T [B,G,3], R [B,G,3,3], thr [B,G,10]}.
        Returns a dict of the per-scene outputs a caller would keep.  collect: optional dict that receives every index
        tensor, forward tensor and gradient of the chain (parity tests; costs two extra query launches per level).
        Cross-step pipelining of the latency-bound sampling chain (it depends on the coordinates only, which a data loader
        has one step ahead): samples = buffers a previous call filled (then this step launches no FPS of its own);
        prefetch = (next_xyz, buffers[, event that next_xyz is ready]): the NEXT step's chain runs on the sampling stream
        beside this step's work."""
        bw = self.backward
        out = {}
        if collect is not None and self.overlap:
            raise ValueError("collect= needs the single-stream schedule (overlap=False)")
        main = torch.cuda.current_stream(self.device) if self.overlap else None

        def sample(cur, npoint):  # furthest_point_sample + gather_operation of one SA module (pointnet2_modules.py:151-158)
            if self.fused_sampling:  # one launch: the FPS kernel holds the coordinates of every pick
                return pu.furthest_point_sample_xyz(cur, npoint)
            inds = pu.furthest_point_sample(cur, npoint)
            return inds, pu.gather_operation(cur.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()

        def collide():
            if self.batched_collision:  # all scenes' occupancy tests in one launch (packed points + offsets)
                return collision_counts_batched(grasps["scene_points"], grasps["T"], grasps["R"], grasps["thr"])
            return torch.stack([collision_counts(grasps["scene_points"][b], grasps["T"][b], grasps["R"][b], grasps["thr"][b])
                                for b in range(len(grasps["scene_points"]))])

        col_done, prefetch_done = None, None
        given = samples is not None
        if given:  # filled by an earlier call that has finished on this stream's timeline: nothing to wait for
            samples = [(i, x, None) for (i, x) in samples]
        else:
            samples = []
        if self.overlap:
            start = torch.cuda.Event()
            start.record(main)
            self._fps_stream.wait_event(start)
            if not given:
                with torch.cuda.stream(self._fps_stream):
                    cur = xyz
                    for (npoint, _, _, _) in SA_SPECS:
                        inds, new_xyz = sample(cur, npoint)
                        ev = torch.cuda.Event()
                        ev.record(self._fps_stream)
                        inds.record_stream(main), new_xyz.record_stream(main)
                        samples.append((inds, new_xyz, ev))
                        cur = new_xyz

        aux_done = None
        pending = []  # (grouped tensor, its gradient): backward_last runs them after every forward, last grouping first
        cur_xyz, level_xyz = xyz, []
        for lvl, (npoint, radius, nsample, c_in) in enumerate(SA_SPECS):
            # ---- SA module (variant A) ----
            if self.overlap or given:
                inds, new_xyz, ev = samples[lvl]
                if ev is not None:
                    main.wait_event(ev)
            else:
                inds, new_xyz = sample(cur_xyz, npoint)
            feats = self.sa_in_feats[lvl]
            grouped, _ = self.sa_groupers[lvl](cur_xyz, new_xyz, feats)
            if bw and feats is not None:
                if self.backward_last:
                    pending.append((grouped, self.sa_grads[lvl]))
                else:
                    grouped.backward(self.sa_grads[lvl])
            if collect is not None:
                collect[f"sa{lvl}_inds"], collect[f"sa{lvl}_xyz"], collect[f"sa{lvl}_grouped"] = inds, new_xyz, grouped.detach()
                collect[f"sa{lvl}_idx"] = pu.ball_query(radius, nsample, cur_xyz, new_xyz)
            if lvl == 0:
                out["sa1_inds"] = inds
            # ---- InvResMLP blocks (variant B) ----
            blocks, c, _, _ = IRM_SPECS[lvl]
            f = self.irm_feats[lvl]
            for _ in range(blocks):
                dp, fj = self.irm_groupers[lvl](new_xyz, new_xyz, f)
                if bw:
                    if self.backward_last:
                        pending.append((fj, self.irm_grads[lvl]))
                    else:
                        fj.backward(self.irm_grads[lvl])
            if collect is not None:
                collect[f"irm{lvl}_dp"], collect[f"irm{lvl}_fj"] = dp, fj.detach()
                collect[f"irm{lvl}_idx"] = gb_group.ball_query(IRM_SPECS[lvl][2], IRM_SPECS[lvl][3], new_xyz, new_xyz)
            cur_xyz = new_xyz
            level_xyz.append(new_xyz)
            if self.overlap and lvl == 0:
                # Host enqueue order = priority: the sampling chain and level 0 of the main stream are on the critical path
                # and were enqueued first; the independent side work follows.  Collision tests: own stream.  The 16
                # cylinder-query crops and the three interpolation chains need only the sampled coordinates: aux stream.
                if grasps is not None:
                    self._col_stream.wait_event(start)
                    with torch.cuda.stream(self._col_stream):
                        out["collision_counts"] = collide()
                        out["collision_counts"].record_stream(main)
                        col_done = torch.cuda.Event()
                        col_done.record(self._col_stream)
                if prefetch is not None:  # the next step's sampling chain: enqueued behind this step's critical path
                    with torch.cuda.stream(self._fps_stream):
                        if len(prefetch) > 2 and prefetch[2] is not None:
                            self._fps_stream.wait_event(prefetch[2])  # the next step's coordinates have arrived
                        self.sampling_chain(prefetch[0], prefetch[1], background=True)
                        prefetch_done = torch.cuda.Event()
                        prefetch_done.record(self._fps_stream)
                self._aux_stream.wait_event(start)
                if samples[1][2] is not None:
                    self._aux_stream.wait_event(samples[1][2])
                with torch.cuda.stream(self._aux_stream):
                    self._crops(xyz, view_rot, samples[1][1], out)
                    if samples[3][2] is not None:
                        self._aux_stream.wait_event(samples[3][2])
                    self._interpolation(xyz, samples[1][1], samples[2][1], samples[3][1], out)
                    for k in ("up_checksum", "crop_checksum"):
                        out[k].record_stream(main)
                    aux_done = torch.cuda.Event()
                    aux_done.record(self._aux_stream)
        sa1_xyz, sa2_xyz, sa3_xyz, sa4_xyz = level_xyz
        out["seed_inds"] = out["sa1_inds"][:, :NUM_SEED]
        for t, g in reversed(pending):  # the order loss.backward() visits the groupings in (drp.py:161-247 forward order reversed)
            t.backward(g)
        pending.clear()
        if not self.overlap:
            self._interpolation(xyz, sa2_xyz, sa3_xyz, sa4_xyz, out, collect)
            self._crops(xyz, view_rot, sa2_xyz, out, collect)
        if aux_done is not None:
            main.wait_event(aux_done)
        if prefetch_done is not None:
            main.wait_event(prefetch_done)
        elif prefetch is not None:  # single-stream schedule
            if len(prefetch) > 2 and prefetch[2] is not None:
                torch.cuda.current_stream(self.device).wait_event(prefetch[2])
            self.sampling_chain(prefetch[0], prefetch[1])
        # ---- collision test ----
        if col_done is not None:
            main.wait_event(col_done)
        elif grasps is not None:
            out["collision_counts"] = collide()
        if bw:
            out["grad_checksum"] = torch.cat([t.grad[:, :4, :64].reshape(-1) for t in
                                              self.sa_in_feats[1:] + self.irm_feats + self.fp_feats]).sum()
            if collect is not None:
                for lvl in range(1, 4):
                    collect[f"sa{lvl}_grad"] = self.sa_in_feats[lvl].grad
                for lvl in range(4):
                    collect[f"irm{lvl}_grad"] = self.irm_feats[lvl].grad
                for i in range(3):
                    collect[f"fp{i}_grad"] = self.fp_feats[i].grad
            for t in self.sa_in_feats[1:] + self.irm_feats + self.fp_feats:
                t.grad = None
        return out


def algorithmic_bytes_per_scene(n=20000, backward=True):
    """HBM bytes one scene's pipeline must move at minimum (SURVEY.md 8d formulas), by op family."""
    by = {"fps": 0, "gather": 0, "ball_query": 0, "cylinder_query": 0, "group_fwd": 0, "group_bwd": 0, "three_nn": 0,
          "interp_fwd": 0, "interp_bwd": 0, "collision": 0}
    cur = n
    for lvl, (m, _, ns, c_in) in enumerate(SA_SPECS):
        by["fps"] += 12 * cur + 4 * m
        by["gather"] += 4 * 3 * cur + 4 * m + 4 * 3 * m
        by["ball_query"] += 12 * cur + 12 * m + 4 * m * ns
        for c in ([3] if c_in == 0 else [3, c_in]):
            by["group_fwd"] += 4 * c * cur + 4 * m * ns + 4 * c * m * ns
        if c_in and backward:
            by["group_bwd"] += 4 * c_in * cur + 4 * m * ns + 4 * c_in * m * ns
        blocks, c, _, nsb = IRM_SPECS[lvl]
        by["ball_query"] += blocks * (24 * m + 4 * m * nsb)
        by["group_fwd"] += blocks * ((4 * 3 * m + 4 * m * nsb + 4 * 3 * m * nsb) + (4 * c * m + 4 * m * nsb + 4 * c * m * nsb))
        if backward:
            by["group_bwd"] += blocks * (4 * c * m + 4 * m * nsb + 4 * c * m * nsb)
        cur = m
    for (nn, mm) in ((512, 256), (1024, 512), (n, 1024)):
        by["three_nn"] += 12 * nn + 12 * mm + 24 * nn
        by["interp_fwd"] += 4 * 256 * mm + 24 * nn + 4 * 256 * nn
        if backward:
            by["interp_bwd"] += 4 * 256 * mm + 24 * nn + 4 * 256 * nn
    # grasp crops: per radius the cloud, seeds and rotations are read once for the four depths
    by["cylinder_query"] += 12 * n + 48 * NUM_SEED + 4 * 16 * NUM_SEED * 64  # one scan for the 4 radii x 4 depths
    by["group_fwd"] += 4 * (4 * 3 * n + 4 * 4 * NUM_SEED * 64 + 4 * 3 * 4 * NUM_SEED * 64)
    by["collision"] += 24 * 5000 + 120 * NUM_GRASP + NUM_GRASP
    return by


def make_view_rotations(batch, seed=0, num_seed=NUM_SEED):
    """Approach frames for the seeds, built like the reference's view templates + batch_viewpoint_params_to_matrix
    (loss_utils.py:15-49): Fibonacci-sphere views, zero in-plane angle."""
    from .scenes import viewpoint_rotations
    rng = np.random.default_rng(seed)
    phi = (math.sqrt(5) - 1) / 2
    i = rng.integers(0, 300, (batch, num_seed))
    z = (2 * i + 1) / 300.0 - 1
    views = np.stack([np.sqrt(1 - z ** 2) * np.cos(2 * i * np.pi * phi), np.sqrt(1 - z ** 2) * np.sin(2 * i * np.pi * phi), z],
                     axis=-1).astype(np.float32)
    return viewpoint_rotations(-views, np.zeros((batch, num_seed), np.float32))
