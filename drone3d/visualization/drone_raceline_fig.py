'''
drone3d/visualization/drone_raceline_fig.py of the reference: the OpenGL / imgui viewer.  Visualisation is out of
scope of the B200 hot path; the class exists so that the reference's scripts run headless to the end.
'''


class DroneRacelineWindow:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs
        if kwargs.get('run', True):
            print('[drone3d shim] DroneRacelineWindow: visualisation is not part of this package (headless run)')

    def run(self):
        return None

    def step(self, *a, **k):
        return False

    def update_projection(self, *a, **k):
        return None
