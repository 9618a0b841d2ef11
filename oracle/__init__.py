"""CPU oracle for the CLIP-prefix language-model step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
path: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing being shipped.

Parity status: PINNED.  ``oracle/validate_against_reference.py`` (run in the
dev container, where ``/root/reference`` exists) checks every function of
``oracle/clip_prefix_lm.py`` against the unmodified reference modules
(``src/models/clipcap.py``, ``src/models/vct0.py``) driving HF ``GPT2LMHeadModel``
and against the golden tensors of ``src/models/vct0_test.py``, then writes the
fixtures under ``tests/golden/`` that travel to the GPU box.  ``oracle/executor_steps.py`` (caption labels,
ensemble scoring) is pinned the same way by ``oracle/validate_executor_steps.py``, which executes the reference's own
statements (read from ``/root/reference`` at run time) on seeded inputs.
"""
