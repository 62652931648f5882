"""GloWGR pieces that sit next to the regression hot path: the LOCO apply step that
produces the ``offset_df`` consumed by ``glow_b200.gwas.linear_regression`` / ``logistic_regression``."""
from .loco import loco_offsets, LabelModel

__all__ = ['loco_offsets', 'LabelModel']
