"""Cuboid bookkeeping and coordinate-volume rotation (host side).

Mirror of `/root/reference/mvn/utils/volumetric.py`: `Cuboid3D` (:44-47; returned from
forward and read only by the reference's visualisation), `get_rotation_matrix` (:87-99)
and `rotate_coord_volume` (:102-114).  cv2 rendering of cuboids is out of scope.
"""
import numpy as np
import torch


class Cuboid3D:
    def __init__(self, position, sides):
        self.position = position
        self.sides = sides


def get_rotation_matrix(axis, theta):
    """Rotation about `axis` by `theta` (Euler-Rodrigues), float64 3x3, same element order as the reference."""
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.sqrt(np.dot(axis, axis))
    a = np.cos(theta / 2.0)
    b, c, d = -axis * np.sin(theta / 2.0)
    aa, bb, cc, dd = a * a, b * b, c * c, d * d
    bc, ad, ac, ab, bd, cd = b * c, a * d, a * c, a * b, b * d, c * d
    return np.array([[aa + bb - cc - dd, 2 * (bc + ad), 2 * (bd - ac)],
                     [2 * (bc - ad), aa + cc - bb - dd, 2 * (cd + ab)],
                     [2 * (bd + ac), 2 * (cd - ab), aa + dd - bb - cc]])


def rotate_coord_volume(coord_volume, theta, axis):
    rot = torch.from_numpy(get_rotation_matrix(axis, theta)).to(coord_volume.device, torch.float)
    return rot.mm(coord_volume.reshape(-1, 3).t()).t().reshape(coord_volume.shape)
