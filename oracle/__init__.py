"""CPU oracle for the per-scan preprocessing hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the timed
CPU baseline - never on the product path (the product path is the C-ABI library under
``autodriver_pointcloud_preprocessor_b200/csrc`` and fails loudly when it is missing).

What it restates (numpy / scipy, one function per reference call site, each citing the
reference file:line it follows):

  * ``pc2``      - sensor_msgs_py ``read_points`` / ``create_cloud`` (un-vendored dependency
                   ``ros2/common_interfaces``, distro unpinned; Humble+ numpy API) and the
                   reference's own ``utils.py:51-133,140-199,423-472`` + ``pp.py:546-625``.
  * ``filters``  - Open3D ``remove_non_finite_points`` / ``transform`` / ``crop`` and
                   ``utils.py:240-301``.
  * ``dedup``    - ``utils.py:509-546`` (numpy / torch / open3d back ends).
  * ``voxel``    - Open3D >=0.18 ``voxel_down_sample`` (``pp.py:509-512``).
  * ``outliers`` - Open3D ``remove_statistical_outliers`` (``pp.py:514-519``) and
                   ``remove_radius_outliers`` (TODO at ``pp.py:37``).
  * ``normals``  - Open3D ``estimate_normals`` hybrid search + analytic 3x3 eigen solver
                   (``pp.py:521-530``).
  * ``ransac``   - Open3D legacy ``segment_plane`` (``pp.py:533-543``).
  * ``concat``   - ``pointcloud_concatenator.py:1-5`` (intent only; semantics defined here).
  * ``pipeline`` - ``pp.py:447-544`` ``preprocess`` stage order.

PARITY PINNING.  The reference has no functional tests, golden vectors or fixtures
(``test/`` holds only ament linters), and Open3D / sensor_msgs_py are neither vendored nor
installable here, so:

  * the stages implemented *inside* the reference (``utils.py`` crop / dedup / unpack /
    field mapping / rgb helpers) are PINNED: ``tests/golden/make_golden.py`` imports the
    reference's ``utils.py`` verbatim under stub ROS modules and a duck-typed Open3D shim,
    runs it on seeded inputs and commits the outputs under ``tests/golden/``; the oracle is
    checked against those vectors;
  * the stages delegated to Open3D / sensor_msgs_py (non-finite, transform, voxel, outliers,
    normals, RANSAC, read_points/create_cloud) are "PARITY UNPINNED": the oracle restates the
    published algorithm (Open3D v0.18/0.19 semantics) and *defines* the choices the
    reference leaves open (RANSAC hypothesis generator, voxel output order, reduction
    order).  Each such choice is written next to the function that makes it.
"""
