def get_layers(*args, **kwargs):  # only reached when Grid(atoms=...) is used
    raise NotImplementedError("ase.geometry.get_layers shim")
