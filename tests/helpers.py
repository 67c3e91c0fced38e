"""Shared helpers of the parity tests: the same robot / scene on both sides, band accounting."""
import numpy as np

BAND = 1e-6  # metres: the reference's libccd tolerance (BASELINE.json north_star)


def device_robot_from_oracle(mg, orobot):
    """Hand the oracle's robot, link for link and joint for joint, to the C ABI."""
    boxes, joints = orobot.describe()
    shapes = [(int(b[0]), b[1:4], b[4:11]) for b in boxes] + list(getattr(orobot, "other_shapes", []))  # capsules / convex links
    jds = [(int(j[0]), int(j[1]), j[3:10], j[10:17], int(j[2]), j[17:20], j[20], j[21]) for j in joints]
    ee = orobot.find_link("end_effector")
    return mg.Robot(orobot.n_links, shapes, jds, orobot.find_link("flying_base"), ee)


def parity_report(gpu_hit, cpu_hit, clearance):
    """SURVEY §8d parity report: mismatches outside the band must be 0; inside they are only counted."""
    gpu_hit, cpu_hit = np.asarray(gpu_hit, bool), np.asarray(cpu_hit, bool)
    in_band = np.abs(clearance) <= BAND
    mism = gpu_hit != cpu_hit
    return dict(n_total=len(gpu_hit), n_mismatch_outside_band=int((mism & ~in_band).sum()), n_in_band=int(in_band.sum()),
                n_mismatch_in_band=int((mism & in_band).sum()))
