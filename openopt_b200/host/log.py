"""Print-based log functions and the package exception (reference: kernel/oologfcn.py:1-34)."""


class OpenOptException(BaseException):
    """Raised by p.err(); a BaseException like the reference's (kernel/oologfcn.py:1)."""

    def __init__(self, msg):
        BaseException.__init__(self, msg)
        self.msg = msg

    def __str__(self):
        return self.msg


class isSolved(BaseException):          # kernel/ooMisc.py:280-281
    pass


class killThread(BaseException):        # kernel/ooMisc.py:282-283
    pass


def oowarn(msg):
    print('OpenOpt Warning: %s' % msg)


def ooerr(msg):
    print('OpenOpt Error: %s' % msg)
    raise OpenOptException(msg)


_printed_once = set()


def ooPWarn(msg):
    if msg not in _printed_once:
        _printed_once.add(msg)
        oowarn(msg)


def ooinfo(msg):
    print('OpenOpt info: %s' % msg)


def oohint(msg):
    print('OpenOpt hint: %s' % msg)


def oodisp(msg):
    print(msg)
