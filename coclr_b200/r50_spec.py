"""Architecture table of ResNet2d3d-50 ("r50"), as data.

Follows backbone/resnet_2d3d.py of the reference: stem conv (5,7,7)/2 + BN + ReLU + max-pool (:138-141), then
[Bottleneck2d x3, Bottleneck2d x4, Bottleneck3d x6, Bottleneck3d x3] with planes 64/128/256/512 and spatial stride
1/2/2/2 (:143-146, 204-208); the first block of every layer carries the 1x1x1 strided downsample conv + BN
(:151-165).  Output [B, 2048, T/2, H/32, W/32].
"""

# (layer name, planes, blocks, spatial stride of the first block, Bottleneck3d?)
R50_LAYERS = [("layer1", 64, 3, 1, False), ("layer2", 128, 4, 2, False),
              ("layer3", 256, 6, 2, True), ("layer4", 512, 3, 2, True)]
R50_FEATURE_SIZE = 2048


def r50_stages(first_channel=3):
    """Ordered stage list for coclr_b200.engine.Graph. Entries:
       ("convbn", conv name, bn name, cin, cout, kernel, stride, padding)     conv -> BN -> ReLU
       ("pool", name, kernel, stride, padding)                                nn.MaxPool3d
       ("bottleneck", name, inplanes, planes, stride, is3d, has_downsample)"""
    st = [("convbn", "conv1", "bn1", first_channel, 64, (5, 7, 7), (2, 2, 2), (2, 3, 3)),
          ("pool", "maxpool", (1, 3, 3), (1, 2, 2), (0, 1, 1))]
    inplanes = 64
    for name, planes, blocks, stride, is3d in R50_LAYERS:
        for i in range(blocks):
            st.append(("bottleneck", "%s.%d" % (name, i), inplanes, planes, stride if i == 0 else 1, is3d, i == 0))
            inplanes = planes * 4
    return st
