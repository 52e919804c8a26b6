"""Stand-in for matplotlib (not installed in this image; test infrastructure): robogym/robot/utils/reach_helper.py:11 imports
pyplot for an optional debug plot that the tests never draw."""
