from .functions import *
