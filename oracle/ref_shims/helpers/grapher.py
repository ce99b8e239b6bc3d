class Grapher(object):
    """No-op grapher (visdom / tensorboardX are not installed); call sites /root/reference/main.py:452-460."""

    def __init__(self, *args, **kwargs):
        pass

    def add_scalar(self, *args, **kwargs):
        pass

    def add_image(self, *args, **kwargs):
        pass

    def add_text(self, *args, **kwargs):
        pass

    def save(self):
        pass

    def close(self):
        pass
