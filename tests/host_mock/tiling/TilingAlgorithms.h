// stand-in for core/tiling/TilingAlgorithms.h:66-116 (TilingAlgorithmBase only)
#pragma once
#include "datastructures/PointBuffer.h"
#include "io/PointsPersistence.h"
#include "process/Tiler.h"
#include "tiling/Sampling.h"
#include <containers/Range.h>
#include <debug/ProgressReporter.h>
#include <taskflow/taskflow.hpp>
#include <util/Definitions.h>
struct TilingAlgorithmBase
{
  TilingAlgorithmBase(SamplingStrategy& sampling_strategy,
                      ProgressReporter* progress_reporter,
                      PointsPersistence& persistence,
                      TilerMetaParameters meta_parameters)
    : _sampling_strategy(sampling_strategy)
    , _progress_reporter(progress_reporter)
    , _persistence(persistence)
    , _meta_parameters(meta_parameters)
  {}
  virtual ~TilingAlgorithmBase() {}
  virtual std::pair<tf::Task, tf::Task> build_execution_graph(util::Range<PointBuffer::PointIterator> points,
                                                              const AABB& bounds,
                                                              uint32_t num_indexing_threads,
                                                              tf::Taskflow& tf) = 0;
  virtual void finalize(const AABB& bounds) {}

protected:
  SamplingStrategy& _sampling_strategy;
  ProgressReporter* _progress_reporter;
  PointsPersistence& _persistence;
  TilerMetaParameters _meta_parameters;
};
