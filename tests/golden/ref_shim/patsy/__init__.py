def dmatrix(*a, **k): raise NotImplementedError("patsy stub")
