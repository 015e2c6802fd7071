"""CPU oracle for the Navier-Stokes step -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``neural-navier-stokes_b200``) never does: it fails loudly without its CUDA library.

* ``oracle.fd``        -- ctypes binding of ``oracle.c`` (chorin_fd + direct_fd, C restatement).
* ``oracle.spectral``  -- numpy restatement of chorin_spectral (needs LAPACK eig/inv).

Parity status: pinned against the reference's own classes run in the build container
(``tests/golden/make_golden.py``); fixtures are committed under ``tests/golden/``.
"""
