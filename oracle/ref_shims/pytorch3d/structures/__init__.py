class Meshes:
    """Plain holder (utils.py:52 builds Meshes(vertices, faces) with a batch of 1)."""

    def __init__(self, verts, faces):
        self.verts, self.faces = verts, faces
