from .clip import *  # noqa: F401,F403  (same surface as the reference's src/clip/__init__.py)
