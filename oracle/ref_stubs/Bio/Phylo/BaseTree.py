class Clade:
    pass
