from . import BaseTree
