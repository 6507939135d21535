"""Status object of the reference (Code/ErrorMessages.f95:31-277): a stack of (state, text) pairs.

Only the boundary convention is mirrored: every host call takes a ``status`` object last and pushes
``setStateTo{Success,Warning,Failure}`` on it from the C-ABI return code and message.
"""
from __future__ import annotations

UndefinedState, SuccessState, WarningState, FailureState = 0, 1, 2, 3
maxNumberOfMessages = 100  # ErrorMessages.f95:25


class ErrorMessage:
    def __init__(self):
        self.currentState = UndefinedState
        self.messages: list[tuple[int, str]] = []

    def _push(self, state, text):
        self.currentState = state
        if text:
            self.messages.append((state, text[:256]))
            del self.messages[:-maxNumberOfMessages]

    def __repr__(self):
        names = {0: "undefined", 1: "success", 2: "warning", 3: "failure"}
        return f"ErrorMessage({names[self.currentState]}, {self.messages[-1][1] if self.messages else ''!r})"


def initializeState(status):
    status.currentState = UndefinedState
    status.messages.clear()


def setStateToSuccess(status, text=""):
    if status is not None and status.currentState != FailureState:
        status._push(SuccessState, text)


def setStateToCompleteSuccess(status, text=""):
    """ErrorMessages.f95:231-248: success that also clears the history."""
    if status is not None:
        status.messages.clear()
        status._push(SuccessState, text)


def setStateToWarning(status, text=""):
    if status is not None and status.currentState != FailureState:
        status._push(WarningState, text)
    elif status is not None and text:
        status.messages.append((WarningState, text[:256]))


def setStateToFailure(status, text=""):
    if status is not None:
        status._push(FailureState, text)


def stateIsSuccess(status):
    return status.currentState == SuccessState


def stateIsWarning(status):
    return status.currentState == WarningState


def stateIsFailure(status):
    return status.currentState == FailureState


def getCurrentMessage(status):
    return status.messages[-1][1] if status.messages else ""


def from_return_code(status, rc, message, where=""):
    """Map a C-ABI return code (0/1/2) + message to the status object."""
    if status is None:
        if rc == 2:
            raise RuntimeError(message or where)
        return
    if rc == 2:
        setStateToFailure(status, message or where)
    elif rc == 1:
        setStateToWarning(status, message or where)
    else:
        setStateToSuccess(status)
