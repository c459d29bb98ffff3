class Axes3D:
    pass
