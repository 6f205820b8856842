"""Backbone factory with the reference's signature (backbone/select_backbone.py:4-16)."""
from .s3dg import S3D


def select_backbone(network, first_channel=3):
    """-> (module, {'feature_size': int}). 's3d' runs on the sm_100a engine; other names of the reference
    ('s3dg', 'r50') are not on the accelerated path yet and raise NotImplementedError like unknown names."""
    param = {'feature_size': 1024}
    if network == 's3d':
        model = S3D(input_channel=first_channel)
    else:
        raise NotImplementedError("backbone %r is not available in coclr_b200 (supported: 's3d')" % (network,))
    return model, param
