"""glow_b200: B200-native drop-in for the linear-regression GWAS hot path of projectglow/glow.

    import glow_b200 as glow
    glow.gwas.linear_regression(genotype_df, phenotype_df, covariate_df, offset_df, ...)

The block loop runs in hand-written sm_100a CUDA behind the C ABI in include/glow_b200.h
(glow_b200/lib/libglow_b200.so, built by ``python -m glow_b200.build``); there is no CPU fallback.
"""
from . import gwas, wgr
from .engine import DeviceGroup, DeviceState, GenotypeBlock
from .sources import ArraySource, BgenSource, GenotypeSource, PlinkBedSource, ResidentSource

__all__ = ['gwas', 'wgr', 'DeviceGroup', 'DeviceState', 'GenotypeBlock', 'ArraySource', 'BgenSource', 'GenotypeSource', 'PlinkBedSource',
           'ResidentSource']
__version__ = '0.1.0'
