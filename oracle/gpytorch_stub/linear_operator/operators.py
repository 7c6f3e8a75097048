class LinearOperator:  # annotation only
    pass
