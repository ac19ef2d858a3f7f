class NoSuchDisplayException(Exception):
    pass
