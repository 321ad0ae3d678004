"""Fake ``glfw`` -- TEST INFRASTRUCTURE ONLY. A headless stand-in for the window library used by
the reference's viewer loop (src/viewer/mujoco_viewer.py:13-32,96-139).  ``window_should_close``
counts calls so that a whole reference script runs for exactly ``STEP_BUDGET`` iterations of
``start_main_loop``; registered key callbacks can be fired synthetically (ball_collision.py needs
one SPACE press to un-pause, src/simulation/ball_collision.py:131-141)."""
PRESS = 1
RELEASE = 0
KEY_SPACE = 32
KEY_BACKSPACE = 259
MOUSE_BUTTON_LEFT = 0
MOUSE_BUTTON_RIGHT = 1

STEP_BUDGET = 0            # set by the runner before executing a script
PRESS_SPACE_AT_START = False
_calls = 0
_key_cb = None


def reset(step_budget, press_space=False):
    global STEP_BUDGET, PRESS_SPACE_AT_START, _calls, _key_cb
    STEP_BUDGET, PRESS_SPACE_AT_START, _calls, _key_cb = int(step_budget), bool(press_space), 0, None


def init():
    return True


def create_window(w, h, title, monitor, share):
    return object()


def make_context_current(window):
    pass


def swap_interval(n):
    pass


def get_cursor_pos(window):
    return (0.0, 0.0)


def terminate():
    pass


def set_key_callback(window, cb):
    global _key_cb
    _key_cb = cb


def set_mouse_button_callback(window, cb):
    pass


def set_cursor_pos_callback(window, cb):
    pass


def set_scroll_callback(window, cb):
    pass


def window_should_close(window):
    global _calls
    if _calls == 0 and PRESS_SPACE_AT_START and _key_cb is not None:
        _key_cb(window, KEY_SPACE, 0, PRESS, 0)
    _calls += 1
    return _calls > STEP_BUDGET


def get_framebuffer_size(window):
    return (4, 4)


def swap_buffers(window):
    pass


def poll_events():
    pass
