class Atoms:  # placeholder: never instantiated on the sr/es/relax path
    pass
