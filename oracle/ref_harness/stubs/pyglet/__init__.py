"""Stub: the reference imports pyglet.canvas.xlib.NoSuchDisplayException unconditionally (envs/car_flag.py:7)."""
