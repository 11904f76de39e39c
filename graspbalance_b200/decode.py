"""Network output -> grasp rows -> collision masks without leaving the device.

`pred_decode` is the drop-in for TrainModel/graspbalance.py:139-192: per scene it picks the best in-plane angle and depth of
every seed, keeps the seeds classified as graspable and emits the [Ns,17] rows graspnetAPI's GraspGroup is built from
(score, width, height, depth, rotation matrix (9), centre (3), object id).  The reference then copies the rows to the host,
wraps them in a GraspGroup and runs the numpy collision detector; `collision_masks` hands the same rows -- still CUDA
tensors, float32 as the reference's -- to ModelFreeCollisionDetector.detect_device, whose arithmetic follows the rows'
dtype exactly as the reference's numpy expressions do.  Same values as the reference functions, evaluated for all scenes
of the batch at once instead of in a Python loop.
"""
import math

import torch

GRASP_MAX_WIDTH = 0.1        # loss_utils.py:6
GRASP_MAX_TOLERANCE = 0.05   # loss_utils.py:7


def batch_viewpoint_params_to_matrix(batch_towards, batch_angle):
    """loss_utils.py:33-49: x axis = approach direction, y axis = (-a_y, a_x, 0) ((0,1,0) when that vanishes), both
    normalised, z = x cross y, then a rotation by the in-plane angle about x.  [N,3], [N] -> [N,3,3]."""
    ax = batch_towards
    zero = torch.zeros_like(ax[:, 0])
    one = torch.ones_like(ax[:, 0])
    ay = torch.stack([-ax[:, 1], ax[:, 0], zero], dim=-1)
    ay[torch.norm(ay, dim=-1) == 0, 1] = 1
    ax = ax / torch.norm(ax, dim=-1, keepdim=True)
    ay = ay / torch.norm(ay, dim=-1, keepdim=True)
    az = torch.cross(ax, ay, dim=-1)
    s, c = torch.sin(batch_angle), torch.cos(batch_angle)
    in_plane = torch.stack([one, zero, zero, zero, c, -s, zero, s, c], dim=-1).reshape(-1, 3, 3)
    return torch.matmul(torch.stack([ax, ay, az], dim=-1), in_plane)


def pred_decode(end_points):
    """graspbalance.py:139-192.  end_points: objectness_score [B,2,Ns], grasp_score_pred / grasp_angle_cls_pred /
    grasp_width_pred / grasp_tolerance_pred [B,A,Ns,D], fp2_xyz [B,Ns,3], grasp_top_view_xyz [B,Ns,3].  Returns a list of B
    tensors [Ns_b,17] on the inputs' device."""
    objectness = end_points['objectness_score'].float()
    score = end_points['grasp_score_pred'].float()
    centre = end_points['fp2_xyz'].float()
    approaching = -end_points['grasp_top_view_xyz'].float()
    angle_cls = end_points['grasp_angle_cls_pred']
    width = torch.clamp(1.2 * end_points['grasp_width_pred'], min=0, max=GRASP_MAX_WIDTH)
    tolerance = end_points['grasp_tolerance_pred']
    B, Ns = centre.shape[0], centre.shape[1]

    best_angle = torch.argmax(angle_cls, 1)                         # [B,Ns,D]
    angle = best_angle.float() / 12 * math.pi
    pick = best_angle.unsqueeze(1)
    score = torch.gather(score, 1, pick).squeeze(1)
    width = torch.gather(width, 1, pick).squeeze(1)
    tolerance = torch.gather(tolerance, 1, pick).squeeze(1)

    best_depth = torch.argmax(score, 2, keepdim=True)               # [B,Ns,1]
    depth = (best_depth.float() + 1) * 0.01
    score = torch.gather(score, 2, best_depth)
    angle = torch.gather(angle, 2, best_depth)
    width = torch.gather(width, 2, best_depth)
    tolerance = torch.gather(tolerance, 2, best_depth)

    graspable = torch.argmax(objectness, 1) == 1                    # [B,Ns]
    score = score * torch.softmax(objectness, dim=1)[:, 1, :].unsqueeze(2)
    score = score * tolerance / GRASP_MAX_TOLERANCE

    rot = batch_viewpoint_params_to_matrix(approaching.reshape(B * Ns, 3), angle.reshape(B * Ns)).reshape(B, Ns, 9)
    rows = torch.cat([score, width, 0.02 * torch.ones_like(score), depth, rot, centre, -1 * torch.ones_like(score)], dim=-1)
    return [rows[b][graspable[b]] for b in range(B)]


def collision_masks(grasp_preds, scene_clouds, voxel_size=0.01, approach_dist=0.05, collision_thresh=0.01):
    """The collision filter the reference applies to decoded grasps (ModelFreeCollisionDetector over each scene's cloud, then
    detect on the scene's GraspGroup), with the rows and the masks staying on the device.  grasp_preds = pred_decode's list;
    scene_clouds = per-scene [N,3] CUDA tensors or numpy arrays.  Returns a list of bool CUDA tensors [Ns_b]."""
    from .collision_detector import ModelFreeCollisionDetector
    out = []
    for rows, cloud in zip(grasp_preds, scene_clouds):
        det = ModelFreeCollisionDetector(cloud, voxel_size=voxel_size, device=rows.device)
        out.append(det.detect_device(rows.contiguous(), approach_dist=approach_dist, collision_thresh=collision_thresh))
    return out
