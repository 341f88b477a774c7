"""Small helpers kept from src/pyclaw/util.py and controller.py."""


class FrameCounter:
    """controller.py (util) frame counter"""

    def __init__(self):
        self.__frame = 0

    def __repr__(self):
        return str(self.__frame)

    def increment(self):
        self.__frame += 1

    def set_counter(self, new_frame_num):
        self.__frame = new_frame_num

    def get_counter(self):
        return self.__frame

    def reset_counter(self):
        self.__frame = 0


def _info_from_argv(argv):
    """util.py:29-55: ``key=value`` command line arguments -> kwargs"""
    args, kwargs = [], {}
    for a in argv[1:]:
        if '=' in a:
            k, v = a.split('=', 1)
            try:
                v = eval(v, {}, {})
            except Exception:
                pass
            kwargs[k] = v
        else:
            args.append(a)
    return args, kwargs


def run_app_from_main(application):
    import sys
    args, kwargs = _info_from_argv(sys.argv)
    return application(*args, **kwargs)
