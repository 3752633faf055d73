from .modules import (KPConv, GroupNorm, UnaryBlock, LastUnaryBlock, ConvBlock, ResidualBlock, maxpool,
                      nearest_upsample, default_kernel_points)
