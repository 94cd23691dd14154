"""pymunk.pygame_util stand-in: rendering is off the hot path (oracle only)."""


class DrawOptions(object):
    def __init__(self, surface=None):
        self.surface = surface
