from .lin_reg import linear_regression

__all__ = ['linear_regression']
