class GaussianBlur(object):          # /root/reference/main.py:384,396 (input pipeline; out of scope)
    def __init__(self, kernel_size, p=0.5):
        self.kernel_size, self.p = kernel_size, p

    def __call__(self, img):
        return img
