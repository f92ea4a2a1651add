// TEST INFRASTRUCTURE.  The reference defines the `Constants` statics in common/motion_planning.cc,
// which needs yaml-cpp (absent offline).  This translation unit defines the same statics and sets them
// the way readAgentConfig (common/motion_planning.cc:54-109) does from the shipped config.yaml
// (r 3, deltat 0.706, LF 2, LB 1, carWidth 2, WB 1, obsRadius 0.8, ...): double expressions assigned to
// float members.  ref_set_constants lets a test use other vehicle values.
#include "common/motion_planning.h"

float Constants::r = 3.0;
float Constants::deltat = 0.706;
float Constants::penaltyTurning = 1.5;
float Constants::penaltyReversing = 2.0;
float Constants::penaltyCOD = 2.0;
float Constants::mapResolution = 2.0;
float Constants::xyResolution = 0;
float Constants::yawResolution = 0;
float Constants::maxClosedSetSize = 1e5;
float Constants::carWidth = 2.0;
float Constants::LF = 2.0;
float Constants::LB = 1.0;
float Constants::WB = 1.0;
float Constants::f2x = 0;
float Constants::r2x = 0;
float Constants::rv = 0;
float Constants::obsRadius = 0.8;
float Constants::constraintWaitTime = 2;
float Constants::speed = 1;
float Constants::t_inc = 0;
double Constants::dubinsShotDistanceSquare = 100;
std::vector<double> Constants::dx, Constants::dy, Constants::dyaw;

extern "C" void ref_set_constants(double r, double deltat, double LF, double LB, double carWidth, double WB,
                                  double obsRadius) {
  Constants::r = r;
  Constants::deltat = deltat;
  Constants::xyResolution = Constants::r * Constants::deltat;
  Constants::yawResolution = Constants::deltat;
  Constants::carWidth = carWidth;
  Constants::LF = LF;
  Constants::LB = LB;
  Constants::WB = WB;
  Constants::f2x = 1 / 4.0 * (3.0 * Constants::LF - Constants::LB);
  Constants::r2x = 1 / 4.0 * (Constants::LF - 3.0 * Constants::LB);
  Constants::rv = 1.0 / 2.0 * pow(pow(Constants::LF + Constants::LB, 2) / 4 + Constants::carWidth * Constants::carWidth, 0.5);
  Constants::obsRadius = obsRadius;
  Constants::speed = 1.0;
  Constants::t_inc = Constants::deltat * Constants::r / Constants::speed;
  Constants::dx = {Constants::r * Constants::deltat, Constants::r * sin(Constants::deltat),
                   Constants::r * sin(Constants::deltat), -Constants::r * Constants::deltat,
                   -Constants::r * sin(Constants::deltat), -Constants::r * sin(Constants::deltat)};
  Constants::dy = {0, -Constants::r * (1 - cos(Constants::deltat)), Constants::r * (1 - cos(Constants::deltat)),
                   0, -Constants::r * (1 - cos(Constants::deltat)), Constants::r * (1 - cos(Constants::deltat))};
  Constants::dyaw = {0, -Constants::deltat, Constants::deltat, 0, Constants::deltat, -Constants::deltat};
}

namespace {
struct Init { Init() { ref_set_constants(3.0, 0.706, 2.0, 1.0, 2.0, 1.0, 0.8); } } init_;
}
