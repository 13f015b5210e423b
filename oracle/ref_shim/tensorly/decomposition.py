def tucker(*a, **k):
    raise NotImplementedError("tensorly.decomposition.tucker (HOSVD/HOOI) is not restated in the test shim")
