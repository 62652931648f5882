from .lin_reg import linear_regression
from .log_reg import logistic_regression

__all__ = ['linear_regression', 'logistic_regression']
