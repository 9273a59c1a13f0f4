"""Drop-in module name: ``from afskmodem import Receiver, Transmitter`` resolves to the B200 core.
Setting ``afskmodem.LOG_LEVEL`` works as in the reference (it forwards to afskmodem_b200.LOG_LEVEL)."""
import sys as _sys

import afskmodem_b200 as _impl

_sys.modules[__name__] = _impl
