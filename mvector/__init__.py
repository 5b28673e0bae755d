"""Import shim: ``import mvector`` from the repository root resolves to the host-side mirror that lives in
``voiceprintrecognition-pytorch_b200/mvector`` (the package directory name is not a valid Python identifier)."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      'voiceprintrecognition-pytorch_b200', 'mvector')
__path__ = [_impl]
with open(_os.path.join(_impl, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, '__init__.py'), 'exec'))
