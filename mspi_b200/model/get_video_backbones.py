"""Motion-encoder registry, mirrors model/get_video_backbones.py:11-31 (same names, same error)."""
from ..backbones.s3d import S3D_features_only
from ..backbones.sf import SlowFast
from ..backbones.X3D import X3D

_MOTION_ENCODERS = ('mvitv2s', 's3d', 'slowfast4x16', 'morphmlps', 'uniformerb', 'videoswins', 'x3dl')
_SUPPORTED = ('s3d', 'x3dl', 'slowfast4x16')


def video_motion_extractor(cfg):
    motion_encoder = None
    if cfg.MODEL.MOTION_ENCODER == 's3d':
        motion_encoder = S3D_features_only(pool=cfg.MODEL.S3D.POOL_STRIDE)
    elif cfg.MODEL.MOTION_ENCODER == 'x3dl':
        motion_encoder = X3D(cfg.MODEL.X3D.PATH_CFG)
    elif cfg.MODEL.MOTION_ENCODER == 'slowfast4x16':
        motion_encoder = SlowFast(cfg.MODEL.SLOWFAST.PATH_CFG)
    if motion_encoder is None:
        raise Exception("Invalid Motion Encoder!")
    return motion_encoder
