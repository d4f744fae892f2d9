"""Motion-encoder registry, mirrors model/get_video_backbones.py:11-31 (same names, same error)."""
from ..backbones.s3d import S3D_features_only

_MOTION_ENCODERS = ('mvitv2s', 's3d', 'slowfast4x16', 'morphmlps', 'uniformerb', 'videoswins', 'x3dl')
_SUPPORTED = ('s3d',)


def video_motion_extractor(cfg):
    motion_encoder = None
    if cfg.MODEL.MOTION_ENCODER == 's3d':
        motion_encoder = S3D_features_only(pool=cfg.MODEL.S3D.POOL_STRIDE)
    if motion_encoder is None:
        raise Exception("Invalid Motion Encoder!")
    return motion_encoder
