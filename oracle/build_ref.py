"""Compile the reference's own Python modules for this path into ``oracle/_ref/`` (git-ignored, travels to the
GPU box with the repo snapshot like the built ``.so``).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference is pure Python, so "compiling the reference" means
byte-compiling its sources *where they lie* under /root/reference into sourceless ``.pyc`` files:

    python -m oracle.build_ref            # writes oracle/_ref/<package>/<module>.refpyc

No reference source text enters the repo; the outputs are CPython bytecode for the interpreter of this image
(the GPU box runs the same image).  The files carry the extension ``.refpyc`` because the gpurun snapshot drops
``*.pyc``; ``add_to_path()`` installs a meta-path finder that loads them with the stock SourcelessFileLoader, which
then serves ``import experiments.run``, ``import experiments.scenarios``, ``import rls...`` - the UNMODIFIED
reference - to

  * tests/test_reference_exec.py (CPU): the reference's ``make_env`` + ``local_obs_*``
    (experiments/scenarios.py:6-63,124-192) executed on tests/_stubs/multiagent and compared with the oracle;
  * tests/test_gpu_reference_loop.py (GPU): the reference's ``run`` / ``run_test`` loops
    (experiments/run.py:11-200) and ``Trainer`` classes (rls/agent/multiagent/*_fix.py) driven against the CUDA
    drop-in;
  * bench.py's CPU arms: the reference's ``ActorNetwork`` (rls/model/ac_network_multi_gumbel.py:24-67).

The third-party ``multiagent`` package the reference imports (experiments/scenarios.py:2-3) is NOT part of the
reference tree and is not produced here; see tests/_stubs/multiagent/__init__.py.
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('MPE_REFERENCE_ROOT', '/root/reference')
OUT = os.path.join(HERE, '_ref')
EXT = '.refpyc'

MODULES = [
    'experiments/scenarios.py',
    'experiments/run.py',
    'experiments/run_BIC.py',
    'rls/arglist.py',
    'rls/utils.py',
    'rls/replay_buffer.py',
    'rls/model/ac_network_multi_gumbel.py',
    'rls/model/ac_network_model_multi_gumbel.py',
    'rls/model/ac_network_multi_gumbel_BIC.py',
    'rls/agent/multiagent/ddpg_gumbel_fix.py',
    'rls/agent/multiagent/model_ddpg_gumbel_fix.py',
    'rls/agent/multiagent/BIC_gumbel_fix.py',
]


def available():
    """True when the compiled reference is importable from oracle/_ref."""
    return all(os.path.exists(os.path.join(OUT, m[:-3] + EXT)) for m in MODULES)


def build(verbose=True):
    """Byte-compile MODULES from REF into OUT.  Returns False (and builds nothing) without the reference tree."""
    if not os.path.isdir(REF):
        return False
    for rel in MODULES:
        src = os.path.join(REF, rel)
        dst = os.path.join(OUT, rel[:-3] + EXT)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: the path shown in tracebacks; file:line citations in this repo are relative to the reference root
        py_compile.compile(src, cfile=dst, dfile='reference/' + rel, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(OUT, 'VERSION'), 'w') as fp:
        fp.write('python %d.%d.%d, %d modules of %s\n' % (sys.version_info[:3] + (len(MODULES), REF)))
    if verbose:
        print('compiled %d reference modules into %s' % (len(MODULES), OUT))
    return True


class _RefFinder(importlib.abc.MetaPathFinder):
    """Serves the packages compiled into oracle/_ref (``experiments``, ``rls``) and nothing else."""

    def find_spec(self, fullname, path=None, target=None):
        rel = os.path.join(OUT, *fullname.split('.'))
        if os.path.isfile(rel + EXT):
            loader = importlib.machinery.SourcelessFileLoader(fullname, rel + EXT)
            return importlib.util.spec_from_file_location(fullname, rel + EXT, loader=loader)
        if os.path.isdir(rel):  # the reference has no __init__.py files: namespace packages
            spec = importlib.machinery.ModuleSpec(fullname, None, is_package=True)
            spec.submodule_search_locations = [rel]
            return spec
        return None


def add_to_path():
    """Make the compiled reference (and nothing else of the reference) importable; raises if it was never built."""
    if not available():
        raise RuntimeError('oracle/_ref is empty: run `python -m oracle.build_ref` where /root/reference exists')
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    return OUT


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
