"""Architecture tables of the backbones on the hot path, as data.

S3D follows backbone/s3dg.py:135-217 of the reference (stage order, kernel/stride/padding of every
conv and pool, SepInception plane counts at :163-164,174-178,191-192).  The engine, the nn.Module
surface (backbone/s3dg.py here) and the parameter naming are all generated from these tables.
"""

# (module name, in_planes, [b0, b1a, b1b, b2a, b2b, b3b])
S3D_INCEPTION = {
    "Mixed_3b": (192, [64, 96, 128, 16, 32, 32]),
    "Mixed_3c": (256, [128, 128, 192, 32, 96, 64]),
    "Mixed_4b": (480, [192, 96, 208, 16, 48, 64]),
    "Mixed_4c": (512, [160, 112, 224, 24, 64, 64]),
    "Mixed_4d": (512, [128, 128, 256, 24, 64, 64]),
    "Mixed_4e": (512, [112, 144, 288, 32, 64, 64]),
    "Mixed_4f": (528, [256, 160, 320, 32, 128, 128]),
    "Mixed_5b": (832, [256, 160, 320, 32, 128, 128]),
    "Mixed_5c": (832, [384, 192, 384, 48, 128, 128]),
}


def s3d_stages(first_channel=3, gating=False):
    """Ordered stage list. Entries:
       ("st", name, cin, cout, k, spatial_stride, temporal_stride, pad)   separable conv (STConv3d)
       ("basic", name, cin, cout)                                          1x1x1 conv (BasicConv3d)
       ("pool", name, kernel, stride, padding)                             nn.MaxPool3d
       ("mixed", name, cin, planes[, gating])                              SepInception (gating: S3D-G, s3dg.py:68-78)
    `block` membership (reference registers blockN aliases) is in S3D_BLOCKS."""
    st = [("st", "Conv_1a", first_channel, 64, 7, 2, 2, 3),
          ("pool", "MaxPool_2a", (1, 3, 3), (1, 2, 2), (0, 1, 1)),
          ("basic", "Conv_2b", 64, 64),
          ("st", "Conv_2c", 64, 192, 3, 1, 1, 1),
          ("pool", "MaxPool_3a", (1, 3, 3), (1, 2, 2), (0, 1, 1))]
    for n in ("Mixed_3b", "Mixed_3c"):
        st.append(("mixed", n) + S3D_INCEPTION[n] + (bool(gating),))
    st.append(("pool", "MaxPool_4a", (3, 3, 3), (2, 2, 2), (1, 1, 1)))
    for n in ("Mixed_4b", "Mixed_4c", "Mixed_4d", "Mixed_4e", "Mixed_4f"):
        st.append(("mixed", n) + S3D_INCEPTION[n] + (bool(gating),))
    st.append(("pool", "MaxPool_5a", (2, 2, 2), (2, 2, 2), (0, 0, 0)))
    for n in ("Mixed_5b", "Mixed_5c"):
        st.append(("mixed", n) + S3D_INCEPTION[n] + (bool(gating),))
    return st


# reference block containers (backbone/s3dg.py:147,154-157,166-169,180-186,194-197)
S3D_BLOCKS = {
    "block1": ["Conv_1a"],
    "block2": ["MaxPool_2a", "Conv_2b", "Conv_2c"],
    "block3": ["MaxPool_3a", "Mixed_3b", "Mixed_3c"],
    "block4": ["MaxPool_4a", "Mixed_4b", "Mixed_4c", "Mixed_4d", "Mixed_4e", "Mixed_4f"],
    "block5": ["MaxPool_5a", "Mixed_5b", "Mixed_5c"],
}
S3D_FEATURE_SIZE = 1024
