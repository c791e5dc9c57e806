def grad_norm(module, norm_type):  # imported by hippie/model.py:7, never called
    raise NotImplementedError
