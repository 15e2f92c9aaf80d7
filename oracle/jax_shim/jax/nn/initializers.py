class Initializer:      # only used as a type annotation by the reference (nn/preconditioner.py:5, 20)
    pass
