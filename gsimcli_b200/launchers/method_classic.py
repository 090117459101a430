# -*- coding: utf-8 -*-
"""The candidate loop of the homogenisation procedure over realization maps
that already exist (the reference's ``gsimcli(..., skip_dss=True)``,
launchers/method_classic.py:44-248; the loop proper is :151-233).

DSS itself (a closed binary driven through parameter files) is out of scope;
what remains of ``gsimcli()`` is glue around the hot path: for every candidate
station, in the given order, split the network (``take_candidate``), load that
candidate's realization files onto the GPU (``GridFiles.load``, K1), detect and
correct against the local PDFs (``detect``: K2c + K3), write the homogenised
series back into the network (``update_station``) -- the next candidates see
the corrected data among their references -- and finally save the
``*_homogenised_data.csv``.  The grid geometry, which the reference reads from
the DSS parameter file, is passed explicitly.
"""
import os

from ..tools import grid as gr
from ..tools import homog as hmg

is_alive = True


def gsimcli(stations_file, stations_header, no_data, stations_order,
            correct_method, detect_prob, detect_flag, detect_save, outfolder,
            purge_sims, dims, first_coord, cells_size, nsim, rad=0,
            correct_skew=None, correct_percentile=None, optional_stats=None,
            print_status=False, skip_dss=True, device=0):
    """Homogenise the stations of ``stations_file`` one candidate at a time.

    Realization files of candidate number ``i`` (position in
    ``stations_order``) are expected where the reference's DSS run leaves them:
    ``<outfolder>/<basename>_dss_map_st<i>_sim.out`` and ``..._sim_<k>.out`` for
    k = 2..nsim (method_classic.py:162, tools/utils.py:104-123), without header
    lines.  Returns ``(homogenised_file, dnumber_list, fnumber_list)`` like the
    reference.
    """
    if not skip_dss:
        raise NotImplementedError(
            'running DSS is outside this package: simulate the maps with the '
            'reference, then call gsimcli(..., skip_dss=True)')
    if isinstance(stations_file, gr.PointSet):
        stations_pset = stations_file
    else:
        stations_pset = gr.PointSet()
        stations_pset.load(stations_file, nd=no_data, header=stations_header)
    dnumber_list, fnumber_list = [], []
    basename = os.path.basename(outfolder)
    for i, station in enumerate(stations_order):
        if not is_alive:
            raise SystemError('process aborted')
        if print_status:
            print('Processing candidate {0} out of {1} with ID {2}.'.format(
                i + 1, len(stations_order), station))
        candidate, references = hmg.take_candidate(stations_pset, station)
        outfile = os.path.join(outfolder, basename + '_dss_map_st' + str(i) + '_sim.out')
        if detect_save:
            references.save(os.path.join(outfolder, basename + '_references_' + str(i)
                                         + '.prn'), header=False)
            candidate.save(os.path.join(outfolder, basename + '_candidate_' + str(i)
                                        + '.prn'), header=True)
        intermediary = os.path.join(outfolder, basename + '_homogenised_' + str(i) + '.prn')
        sim_maps = gr.GridFiles()
        sim_maps.load(outfile, nsim, dims, first_coord, cells_size, no_data,
                      headerin=0, device=device)
        homogenised, detected_number, filled_number = hmg.detect(
            grids=sim_maps, obs_file=candidate, method=correct_method,
            prob=detect_prob, flag=detect_flag, save=detect_save,
            outfile=intermediary, header=True, skewness=correct_skew, rad=rad,
            percentile=correct_percentile, optional_stats=optional_stats)
        if print_status:
            print('Inhomogeneities detected: {0}'.format(detected_number))
        dnumber_list.append(detected_number)
        fnumber_list.append(filled_number)
        stations_pset = hmg.update_station(stations_pset, homogenised)
        if purge_sims:
            sim_maps.purge()
        else:
            sim_maps.dump()
    homogenised_file = os.path.join(outfolder, basename + '_homogenised_data.csv')
    hmg.save_output(pset_file=stations_pset, outfile=homogenised_file,
                    fformat='gsimcli', header=True, save_stations=True)
    return homogenised_file, dnumber_list, fnumber_list
