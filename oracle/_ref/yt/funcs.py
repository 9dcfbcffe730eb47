import logging
mylog = logging.getLogger("yt")
